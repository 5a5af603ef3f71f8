#!/bin/bash
# final visit without the ncu --set full captures: smoke, full GPU suite, default bench line, ncu launch list
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "== pytest -m gpu (full)"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee $OUT/r2ae_pytest.log
echo "== bench (default, full line)"; timeout 900 python bench.py > $OUT/r2ae_bench_1gpu.json 2> $OUT/r2ae_bench_1gpu.err; tail -2 $OUT/r2ae_bench_1gpu.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:'pos_compact|onesweep|hist_kernel|uniq_|bucket_|slice_|unit_rank|pscan_|rank_kernel|pos_sort|pos_gather|rank_scan|export_pos|head_kernel' \
  --csv --log-file $OUT/r2ae_launches.csv \
  python bench.py --images 1500 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-extra > $OUT/r2ae_launches_bench.log 2>&1
echo "== done"
