#!/bin/bash
# round 2, visit X (2 GPUs): f-4 SyncBN tests (incl. the 2-rank NCCL case) + kernel bandwidth
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_syncbn.py tests/test_gpu_resize.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 | tee $OUT/r2x_tests.log
echo "== bandwidth"; timeout 600 python tools/bench_syncbn.py 2>&1 | tail -1 | tee $OUT/r2x_syncbn.json
echo "== done"
