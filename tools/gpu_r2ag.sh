#!/bin/bash
# round 2, visit AB (1 GPU): loss forward with fewer instructions per class -- tests + the extras block of the bench
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_loss.py tests/test_gpu_deeplab.py tests/test_gpu_configs.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 | tee $OUT/r2ag_tests.log
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --steps 3 --no-e2e > $OUT/r2ag_bench.json 2> $OUT/r2ag_bench.err; tail -2 $OUT/r2ag_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ag_bench.json'))
print(d['ms_per_step'], d['value'])
for e in d['roofline'].get('extra', []): print(e.get('name'), round(e.get('ms', 0), 3), round(e.get('frac', 0) or 0, 3))
PY
echo "== done"
