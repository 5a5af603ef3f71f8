#!/bin/bash
# round 2, visit Y (1 GPU): positives gathered inside the head -- tests, A/B bench
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_rank.py tests/test_gpu_head.py tests/test_gpu_anomaly.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 | tee $OUT/r2y_tests.log
for flag in "" "--no-fused-gather" "" "--no-fused-gather"; do
  tag=fused; [ -n "$flag" ] && tag=separate
  echo "== bench $tag"
  timeout 600 python bench.py --no-cpu-baseline --steps 3 --no-e2e --no-extra $flag > $OUT/r2y_bench_$tag.json 2> $OUT/r2y_bench_$tag.err; tail -2 $OUT/r2y_bench_$tag.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2y_bench_$tag.json'))
print(d['ms_per_step'], d['value'], [ (s['stage'][:30], round(s['ms_per_step'],2)) for s in d['roofline']['stages']], d['roofline']['frac'])
print(d['results'])
PY
done
echo "== done"
