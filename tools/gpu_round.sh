#!/bin/bash
# One GPU-box visit: parity tests, the bench line, A/B runs of the tuning knobs, the ncu launch list and one
# `--set full` capture of the hot kernels.  Everything lands in gpurun_out/ (merged back by gpurun).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round.sh r1d'
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1

echo "== pytest -m gpu" | tee $OUT/${TAG}_pytest.log
timeout 420 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log

echo "== bench (default)"
timeout 330 python bench.py > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench_1gpu.err
tail -c 600 $OUT/${TAG}_bench_1gpu.json

echo "== bench A/B: previous formulation (per-class head, per-key scan, regenerated pooled keys)"
DML_SCAN_BITS=0 DML_HEAD_LEAN=0 timeout 200 python bench.py --pooled-keys regenerate --no-e2e --no-cpu-baseline --steps 3 \
  > $OUT/${TAG}_bench_ab_previous.json 2> $OUT/${TAG}_bench_ab_previous.err

echo "== ncu launch list"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --images 100 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/${TAG}_launches_bench.log 2>&1

echo "== bench A/B: 4 pixels per thread in the head"
DML_HEAD_VEC=4 timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 3 \
  > $OUT/${TAG}_bench_ab_vec4.json 2> $OUT/${TAG}_bench_ab_vec4.err

echo "== pytest with the previous formulations (fallback knobs)" | tee $OUT/${TAG}_pytest_fallback.log
DML_SCAN_BITS=0 DML_HEAD_LEAN=0 timeout 300 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_head.py tests/test_gpu_anomaly.py \
  -m gpu -q -p no:cacheprovider 2>&1 | tail -15 >> $OUT/${TAG}_pytest_fallback.log
tail -2 $OUT/${TAG}_pytest_fallback.log

echo "== ncu --set full: head, key-gen, onesweep, scan kernels (first launches of one 50-image chunk)"
timeout 420 ncu --set full --clock-control none --import-source on \
  -k regex:'head_kernel|keygen_kernel|onesweep_kernel|scan_agg_kernel|scan_apply_kernel' -c 8 -f -o $OUT/${TAG}_full \
  python bench.py --images 50 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/${TAG}_full_bench.log 2>&1
ls -la $OUT | tail -20
echo "== done"
