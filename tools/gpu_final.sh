#!/bin/bash
# round 2, final visit: full GPU suite; default bench; ncu of loss / class sums / partition
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== pytest -m gpu (full)"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12 | tee $OUT/r2ae_pytest.log
echo "== bench (default, full line)"; timeout 900 python bench.py > $OUT/r2ae_bench_1gpu.json 2> $OUT/r2ae_bench_1gpu.err; tail -3 $OUT/r2ae_bench_1gpu.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ae_bench_1gpu.json'))
print(d['ms_per_step'], d['value'], [ (s['stage'][:30], round(s['ms_per_step'],2)) for s in d['roofline']['stages']])
print(d.get('pooled_verified'), d['cpu_baseline'], d['e2e']['value'])
for k,v in d['roofline'].get('extra',{}).items(): print(k, round(v['ms'],3), round(v['frac_of_hbm_peak'],3))
PY
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:'pos_compact|onesweep|hist_kernel|uniq_|bucket_|slice_|unit_rank|pscan_|rank_kernel|pos_sort|pos_gather|rank_scan|export_pos|head_kernel' \
  --csv --log-file $OUT/r2ae_launches.csv \
  python bench.py --images 1500 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-extra > $OUT/r2ae_launches_bench.log 2>&1
python tools/launch_summary.py $OUT/r2ae_launches.csv --last 26 2>/dev/null | tail -50
echo "== ncu full: loss / class sums"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'loss_kernel|class_sums' -c 6 -f -o $OUT/r2ae_loss_protos \
  python tools/bench_loss_protos.py --iters 1 > $OUT/r2ae_loss_protos.log 2>&1
python tools/ncu_summary.py $OUT/r2ae_loss_protos.ncu-rep > $OUT/r2ae_ncu_loss_protos_summary.txt; head -60 $OUT/r2ae_ncu_loss_protos_summary.txt
echo "== ncu full: pooled kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bucket_count|bucket_scatter|unit_rank' -c 3 -f -o $OUT/r2ae_pooled \
  python bench.py --images 592 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-extra > $OUT/r2ae_pooled_bench.log 2>&1
python tools/ncu_summary.py $OUT/r2ae_pooled.ncu-rep | tee $OUT/r2ae_ncu_pooled_summary.txt
echo "== ncu full: f-4 kernels (SyncBN, resize)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'bn_stats|bn_apply|bn_bwd|resize_norm' -c 8 -f -o $OUT/r2ae_f4 \
  python tools/bench_f4.py > $OUT/r2ae_f4.log 2>&1
python tools/ncu_summary.py $OUT/r2ae_f4.ncu-rep > $OUT/r2ae_ncu_f4_summary.txt; head -80 $OUT/r2ae_ncu_f4_summary.txt
echo "== done"
