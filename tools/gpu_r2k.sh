#!/bin/bash
# round 2, visit K (2 GPUs): NCCL tests incl. mode=rank, bench --gpus 2 (weak + strong + config 5 + self-check), f-2 test
set -u
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,index --format=csv
echo "== tests (2 GPUs)"; timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_conv_head.py tests/test_gpu_loss.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 | tee $OUT/r2k_tests_2gpu.log
echo "== bench 2 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/r2k_bench_2gpu.json 2> $OUT/r2k_bench_2gpu.err
tail -3 $OUT/r2k_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench_2gpu.json'))
print("ms/step", d['ms_per_step'], "value", d['value'])
for s in d['roofline']['stages']: print(s['stage'][:50], round(s['ms_per_step'],2), s.get('phases_ms_last_step_rank0'))
print("verified", d.get('pooled_verified'), d.get('pooled_verify'))
print("strong", d.get('strong_scaling'))
print("config5", d.get('config5_metric_sweep'))
print("e2e", d['e2e'])
print(d['results'])
PY
echo "== done"
