#!/bin/bash
# round 2, visit AF (1 GPU): ncu --set full of the class-sum kernels (coherent + iid labels) and of the partition kernels
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'class_sums' -c 6 -f -o $OUT/r2af_class_sums \
  python tools/bench_loss_protos.py --iters 1 > $OUT/r2af_class_sums.log 2>&1
python tools/ncu_summary.py $OUT/r2af_class_sums.ncu-rep > $OUT/r2af_ncu_class_sums_summary.txt; cat $OUT/r2af_ncu_class_sums_summary.txt | head -90
python tools/ncu_lines.py $OUT/r2af_class_sums.ncu-rep class_sums_lane 2>/dev/null | head -40 > $OUT/r2af_class_sums_lines.txt; cat $OUT/r2af_class_sums_lines.txt
echo "== done"
