#!/bin/bash
# round 2, visit C: bucket-sorted positives; ncu --set full of the rank kernels
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== rank tests"; timeout 600 python -m pytest tests/test_gpu_rank.py -m gpu -q -p no:cacheprovider 2>&1 | tail -30 | tee $OUT/r2c_rank_tests.log
echo "== bench rank"; timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 > $OUT/r2c_bench_rank.json 2> $OUT/r2c_bench_rank.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench_rank.json'))
print(d['ms_per_step'], [ (s['stage'][:30], round(s['ms_per_step'],2)) for s in d['roofline']['stages']])
PY
echo "== ncu full"
timeout 420 ncu --set full --clock-control none --import-source on -k regex:'rank_kernel|pos_sort_kernel|pos_gather_kernel|rank_scan_kernel' -c 4 -f -o $OUT/r2c_full \
  python bench.py --images 74 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-pooled > $OUT/r2c_full_bench.log 2>&1
python tools/ncu_summary.py $OUT/r2c_full.ncu-rep | tee $OUT/r2c_ncu_summary.txt
echo "== done"
