#!/bin/bash
# round 2, visit V (N GPUs): PositiveExchange -- NCCL tests, bench A/B (overlapped vs bulk exchange of the positives)
set -u
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 | tee $OUT/r2v_tests_${N}gpu.log
for flag in "" "--no-overlap-exchange"; do
  tag=overlap; [ -n "$flag" ] && tag=bulk
  echo "== bench $tag"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --config5-images 0 $flag > $OUT/r2v_bench_${N}gpu_$tag.json 2> $OUT/r2v_bench_${N}gpu_$tag.err
  tail -2 $OUT/r2v_bench_${N}gpu_$tag.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2v_bench_${N}gpu_$tag.json'))
print("ms/step", d['ms_per_step'], "value", d['value'])
for s in d['roofline']['stages']: print(s['stage'][:40], round(s['ms_per_step'],2), s.get('phases_ms_last_step_rank0'))
print("verified", d.get('pooled_verified'), d.get('pooled_verify'))
print("strong", d.get('strong_scaling'))
print(d['results'])
PY
done
echo "== done"
