#!/bin/bash
# round 2, visit O (1 GPU): CUB yardstick next to this repo's sort; tests; bench with the reverted short cuts + pos_gather prefetch
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_rank.py tests/test_gpu_conv_head.py tests/test_gpu_loss.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 | tee $OUT/r2o_tests.log
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --steps 3 --no-e2e > $OUT/r2o_bench.json 2> $OUT/r2o_bench.err; tail -3 $OUT/r2o_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2o_bench.json'))
print(d['ms_per_step'], d['value'], [ (s['stage'][:30], round(s['ms_per_step'],2)) for s in d['roofline']['stages']])
print(d.get('pooled_verified'))
PY
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:'bucket_|unit_rank|rank_kernel|pos_sort|pos_gather|rank_scan|export_pos|head_kernel' \
  --csv --log-file $OUT/r2o_launches.csv \
  python bench.py --images 296 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-extra > $OUT/r2o_launches_bench.log 2>&1
python tools/launch_summary.py $OUT/r2o_launches.csv --last 14 2>/dev/null | tail -34
echo "== done"
