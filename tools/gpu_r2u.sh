#!/bin/bash
# round 2, visit U (1 GPU): A/B of the two unit_rank chunk walks in the same visit (strong / full / weak regimes)
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for mode in chunk flat chunk flat; do
  echo "== DML_UNIT_RANK=$mode"
  DML_UNIT_RANK=$mode timeout 600 python tools/strong_regime.py --shards 187,375,750,1500 2>&1 | grep shard
  DML_UNIT_RANK=$mode timeout 600 python tools/strong_regime.py --positive-passes 8 --shards 1500 2>&1 | grep shard
done | tee $OUT/r2u_ab.log
echo "== done"
