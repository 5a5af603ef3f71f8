#!/bin/bash
# round 2, visit E: where does the pooled rank stage spend its time; division check; updated rank kernels
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== check_div"; timeout 120 tools/bin/check_div | tee $OUT/r2e_check_div.json
echo "== rank tests"; timeout 900 python -m pytest tests/test_gpu_rank.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -20 | tee $OUT/r2e_rank_tests.log
echo "== ncu launch list (pooled kernels only)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:'pos_compact|onesweep|hist_kernel|hist_scan|uniq_|scan_u32|bucket_|unit_rank|pscan_|rank_kernel|pos_sort|pos_gather|rank_scan|head_kernel' \
  --csv --log-file $OUT/r2e_launches.csv \
  python bench.py --images 1500 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-extra > $OUT/r2e_launches_bench.log 2>&1
python tools/launch_summary.py $OUT/r2e_launches.csv --last 40 2>/dev/null | tail -70
echo "== done"
