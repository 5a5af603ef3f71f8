#!/bin/bash
# round 2, visit T (1 GPU): flattened unit_rank
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_rank.py tests/test_gpu_distributed.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 | tee $OUT/r2t_tests.log
echo "== strong regime"; timeout 600 python tools/strong_regime.py 2>&1 | tee $OUT/r2t_strong_regime.log
echo "== weak regime"; timeout 600 python tools/strong_regime.py --positive-passes 8 --shards 1500 2>&1 | tee $OUT/r2t_weak_regime.log
echo "== launch list at 187 images"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2t_launches.csv -k regex:"bucket|unit_rank|slice" python tools/strong_regime.py --shards 187,1500 --reps 1 > $OUT/r2t_ncu.log 2>&1
python tools/launch_summary.py $OUT/r2t_launches.csv 2>&1 | grep -E "unit_rank|scatter_kernel  " | head -20
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --steps 3 --no-e2e --no-extra > $OUT/r2t_bench.json 2> $OUT/r2t_bench.err; tail -3 $OUT/r2t_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2t_bench.json'))
print(d['ms_per_step'], d['value'], [ (s['stage'][:30], round(s['ms_per_step'],2)) for s in d['roofline']['stages']], d.get('pooled_verified'))
PY
echo "== done"
