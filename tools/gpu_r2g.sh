#!/bin/bash
# round 2, visit H: warp-aggregated scatter, new class_sums kernel
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_rank.py tests/test_gpu_loss.py tests/test_gpu_configs.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -20 | tee $OUT/r2h_tests.log
echo "== bench rank"; timeout 600 python bench.py --no-cpu-baseline --steps 3 --no-e2e > $OUT/r2h_bench_rank.json 2> $OUT/r2h_bench_rank.err; tail -5 $OUT/r2h_bench_rank.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench_rank.json'))
print(d['ms_per_step'], d['value'], [ (s['stage'][:30], round(s['ms_per_step'],2)) for s in d['roofline']['stages']])
print(d.get('pooled_verified'))
for k,v in d['roofline'].get('extra',{}).items(): print(k, round(v['ms'],3), round(v['frac_of_hbm_peak'],3))
PY
echo "== ncu launch list (pooled kernels only)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:'pos_compact|onesweep|hist_kernel|uniq_|bucket_|unit_rank|pscan_|rank_kernel|pos_sort|pos_gather|rank_scan|export_pos' \
  --csv --log-file $OUT/r2h_launches.csv \
  python bench.py --images 1500 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-extra > $OUT/r2h_launches_bench.log 2>&1
python tools/launch_summary.py $OUT/r2h_launches.csv --last 24 2>/dev/null | tail -48
echo "== done"
