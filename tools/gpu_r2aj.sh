#!/bin/bash
# round 2, visit AJ (1 GPU): ncu --set full of head_kernel at the bench's launch shape (74 images) -> DRAM traffic per pixel
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'head_kernel' -c 2 -f -o $OUT/r2aj_head \
  python bench.py --images 148 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-extra --no-pooled > $OUT/r2aj_head.log 2>&1
python tools/ncu_summary.py $OUT/r2aj_head.ncu-rep | tee $OUT/r2aj_ncu_head_summary.txt
echo "== done"
