#!/bin/bash
# round 2, multi-GPU visit: bench.py --gpus N via torchrun (weak scaling + strong + config 5 + self-check + e2e)
set -u
N=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nvidia-smi topo -m > $OUT/r2al_topo_${N}gpu.txt 2>&1
cat /sys/devices/system/node/online; nproc
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/r2al_bench_${N}gpu.json 2> $OUT/r2al_bench_${N}gpu.err
tail -3 $OUT/r2al_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2al_bench_${N}gpu.json'))
print("ms/step", d['ms_per_step'], "value", d['value'])
for s in d['roofline']['stages']: print(s['stage'][:50], round(s['ms_per_step'],2), s.get('phases_ms_last_step_rank0'))
print("verified", d.get('pooled_verified'), d.get('pooled_verify'))
print("strong", d.get('strong_scaling'))
print("config5", d.get('config5_metric_sweep'))
print("e2e", d['e2e'])
print(d['results'])
PY
echo "== done"
