#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_conv_head.py tests/test_gpu_rank.py tests/test_gpu_multiscale.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 | tee $OUT/r2p_tests.log
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --steps 3 --no-e2e --no-extra > $OUT/r2p_bench.json 2> $OUT/r2p_bench.err; tail -3 $OUT/r2p_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p_bench.json'))
print(d['ms_per_step'], d['value'], [ (s['stage'][:30], round(s['ms_per_step'],2)) for s in d['roofline']['stages']])
PY
echo "== done"
