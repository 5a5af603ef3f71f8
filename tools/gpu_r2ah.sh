#!/bin/bash
# round 2, visit AH (1 GPU): per-image metric kernels on a second stream next to the following head (A/B)
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for flag in "--metric-stream" "" "--metric-stream" ""; do
  tag=two; [ -z "$flag" ] && tag=one
  echo "== bench $tag stream(s)"
  timeout 600 python bench.py --no-cpu-baseline --steps 3 --no-e2e --no-extra $flag > $OUT/r2ah_bench_$tag.json 2> $OUT/r2ah_bench_$tag.err; tail -1 $OUT/r2ah_bench_$tag.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2ah_bench_$tag.json'))
print(d['ms_per_step'], d['value'], [ (s['stage'][:30], round(s['ms_per_step'],2)) for s in d['roofline']['stages']], d['roofline']['frac'])
print(d['results'])
PY
done
echo "== done"
