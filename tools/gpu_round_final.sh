#!/bin/bash
# Final 1-GPU visit of the round: parity tests, smoke, the bench line, the ncu launch list and a `--set full` capture.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/gpu_round_final.sh r1f'
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu" | tee $OUT/${TAG}_pytest.log
timeout 420 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
echo "== smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
echo "== bench (default)"
timeout 330 python bench.py > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench_1gpu.err
tail -c 300 $OUT/${TAG}_bench_1gpu.json
echo "== bench --pooled-keys regenerate (separate pooled key generation + hist_kernel)"
timeout 200 python bench.py --pooled-keys regenerate --no-e2e --no-cpu-baseline --steps 3 \
  > $OUT/${TAG}_bench_ab_regenerate.json 2> $OUT/${TAG}_bench_ab_regenerate.err
echo "== ncu launch list"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --images 100 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/${TAG}_launches_bench.log 2>&1
echo "== ncu --set full"
timeout 300 ncu --set full --clock-control none --import-source on \
  -k regex:'head_kernel|keygen_kernel|onesweep_kernel|scan_agg_kernel|scan_apply_kernel' -c 8 -f -o $OUT/${TAG}_full \
  python bench.py --images 50 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/${TAG}_full_bench.log 2>&1
echo "== done"
