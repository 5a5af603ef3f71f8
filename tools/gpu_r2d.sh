#!/bin/bash
# round 2, visit D: pooled minority-rank path: tests, bench
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== rank tests"; timeout 900 python -m pytest tests/test_gpu_rank.py tests/test_gpu_distributed.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -40 | tee $OUT/r2d_rank_tests.log
echo "== bench rank"; timeout 600 python bench.py --no-cpu-baseline --steps 3 > $OUT/r2d_bench_rank.json 2> $OUT/r2d_bench_rank.err; tail -5 $OUT/r2d_bench_rank.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench_rank.json'))
print(d['ms_per_step'], d['value'], [ (s['stage'][:30], round(s['ms_per_step'],2)) for s in d['roofline']['stages']])
print(d.get('pooled_verified'), d.get('pooled_verify'))
print(json.dumps(d['roofline'].get('extra'), indent=1))
print(d['e2e'])
print(d['results'])
PY
echo "== ncu launch list (rank)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2d_launches.csv \
  python bench.py --images 296 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > $OUT/r2d_launches_bench.log 2>&1
python tools/launch_summary.py $OUT/r2d_launches.csv --last 60 2>/dev/null | tail -90
echo "== done"
