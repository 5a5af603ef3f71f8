#!/bin/bash
# round 2, visit AK (1 GPU): ncu --set full of the per-image rank pipeline at the bench's launch shape (74 images)
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rank_kernel|pos_gather|pos_sort|rank_scan|export_pos' -c 5 -f -o $OUT/r2ak_rank \
  python bench.py --images 74 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-extra --no-pooled > $OUT/r2ak_rank.log 2>&1
python tools/ncu_summary.py $OUT/r2ak_rank.ncu-rep | tee $OUT/r2ak_ncu_rank_summary.txt
python tools/ncu_lines.py $OUT/r2ak_rank.ncu-rep rank_kernel 2>/dev/null | head -28 | tee $OUT/r2ak_rank_kernel_lines.txt
echo "== done"
