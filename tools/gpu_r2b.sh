#!/bin/bash
# round 2, visit B: minority-rank path -- parity tests, bench A/B rank vs sort, launch list
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== rank tests"; timeout 600 python -m pytest tests/test_gpu_rank.py tests/test_gpu_fullshape.py -m gpu -q -p no:cacheprovider 2>&1 | tail -40 | tee $OUT/r2b_rank_tests.log
echo "== bench rank"; timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 > $OUT/r2b_bench_rank.json 2> $OUT/r2b_bench_rank.err; tail -c 1500 $OUT/r2b_bench_rank.json; tail -5 $OUT/r2b_bench_rank.err
echo "== bench sort"; timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 --metric-method sort > $OUT/r2b_bench_sort.json 2> $OUT/r2b_bench_sort.err; tail -c 600 $OUT/r2b_bench_sort.json
echo "== ncu launch list (rank)"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2b_launches.csv \
  python bench.py --images 148 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/r2b_launches_bench.log 2>&1
python tools/launch_summary.py $OUT/r2b_launches.csv 2>/dev/null | tail -40
echo "== done"
