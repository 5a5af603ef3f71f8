#!/bin/bash
# round 2, visit Q (1 GPU): new rank-path tests; pooled stage in the strong-scaling regime (shard of the negatives vs global positives)
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_rank.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 | tee $OUT/r2q_tests.log
echo "== strong regime"; timeout 600 python tools/strong_regime.py 2>&1 | tee $OUT/r2q_strong_regime.log
echo "== launch list at 187 images"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2q_launches.csv -k regex:"bucket|unit_rank|slice|uniq|pscan" python tools/strong_regime.py --shards 187 --reps 1 > $OUT/r2q_ncu.log 2>&1
python tools/launch_summary.py $OUT/r2q_launches.csv 2>&1 | tail -40
echo "== done"
