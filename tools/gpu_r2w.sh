#!/bin/bash
# round 2, visit W (1 GPU): f-4 input side (GPU resize + normalise) tests + timing against PIL on the host
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_resize.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 | tee $OUT/r2w_tests.log
echo "== timing"; timeout 600 python tools/bench_resize.py 2>&1 | tee $OUT/r2w_resize.json
echo "== done"
