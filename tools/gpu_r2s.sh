#!/bin/bash
# round 2, visit S (1 GPU): pooled stage emulated in the 8-GPU regimes on one GPU (strong: 1/8 shard vs all positives; weak: full shard vs 8x positives)
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== weak regime (positives of 8 x 1500 images)"
timeout 900 python tools/strong_regime.py --positive-passes 8 --shards 1500 2>&1 | tee $OUT/r2s_weak_regime.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2s_launches_weak.csv -k regex:"bucket|unit_rank|slice" python tools/strong_regime.py --positive-passes 8 --shards 1500 --reps 1 > $OUT/r2s_ncu_weak.log 2>&1
python tools/launch_summary.py $OUT/r2s_launches_weak.csv 2>&1 | tail -9
echo "== ncu full, strong regime: unit_rank + scatter"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"unit_rank|bucket_scatter" -c 2 -o $OUT/r2s_strong python tools/strong_regime.py --shards 187 --reps 1 > $OUT/r2s_ncu_strong.log 2>&1
python tools/ncu_summary.py $OUT/r2s_strong.ncu-rep 2>&1 | tee $OUT/r2s_ncu_strong_summary.txt | head -60
echo "== ncu full, weak regime: unit_rank + scatter"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"unit_rank|bucket_scatter|bucket_count" -c 3 -o $OUT/r2s_weak python tools/strong_regime.py --positive-passes 8 --shards 1500 --reps 1 > $OUT/r2s_ncu_weak2.log 2>&1
python tools/ncu_summary.py $OUT/r2s_weak.ncu-rep 2>&1 | tee $OUT/r2s_ncu_weak_summary.txt | head -60
echo "== done"
