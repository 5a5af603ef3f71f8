#!/bin/bash
# round 2, visit R (1 GPU): unit_rank chunk descriptors in smem + batched plan loads
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_rank.py tests/test_gpu_distributed.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 | tee $OUT/r2r_tests.log
echo "== strong regime"; timeout 600 python tools/strong_regime.py 2>&1 | tee $OUT/r2r_strong_regime.log
echo "== launch list at 187 images"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2r_launches.csv -k regex:"bucket|unit_rank|slice" python tools/strong_regime.py --shards 187 --reps 1 > $OUT/r2r_ncu.log 2>&1
python tools/launch_summary.py $OUT/r2r_launches.csv 2>&1 | tail -12
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --steps 3 --no-e2e --no-extra > $OUT/r2r_bench.json 2> $OUT/r2r_bench.err; tail -3 $OUT/r2r_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2r_bench.json'))
print(d['ms_per_step'], d['value'], [ (s['stage'][:30], round(s['ms_per_step'],2)) for s in d['roofline']['stages']], d.get('pooled_verified'))
PY
echo "== done"
