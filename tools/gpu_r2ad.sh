#!/bin/bash
# round 2, visit AD (1 GPU): fast mix coefficient in rank_kernel -- tests + bench
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_rank.py tests/test_gpu_metrics.py tests/test_gpu_anomaly.py tests/test_gpu_baseline.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 | tee $OUT/r2ad_tests.log
for i in 1 2; do
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --steps 3 --no-e2e --no-extra > $OUT/r2ad_bench.json 2> $OUT/r2ad_bench.err; tail -1 $OUT/r2ad_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ad_bench.json'))
print(d['ms_per_step'], d['value'], [ (s['stage'][:30], round(s['ms_per_step'],2)) for s in d['roofline']['stages']], d['roofline']['frac'])
PY
done
echo "== done"
