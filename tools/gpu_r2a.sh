#!/bin/bash
# round 2, visit A: primitive micro-benchmarks for the minority-rank design, the new full-shape parity tests, the GPU suite
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/r2a_gpu.txt 2>&1
echo "== microbench"; timeout 120 tools/bin/microbench_rank | tee $OUT/r2a_microbench.jsonl
echo "== full-shape parity"; timeout 600 python -m pytest tests/test_gpu_fullshape.py -m gpu -q -p no:cacheprovider 2>&1 | tail -40 | tee $OUT/r2a_fullshape.log
cat $OUT/fullshape_parity.json
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -30 | tee $OUT/r2a_pytest.log
echo "== done"
