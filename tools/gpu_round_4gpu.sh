#!/bin/bash
# Four-GPU visit: the N=4 bench line only (device-resident timing; the end-to-end leg is PCIe-bound per GPU and is
# measured by the round-end scaling run).
#   /usr/local/graft/bin/gpurun --gpus 4 --timeout 150 -- 'bash tools/gpu_round_4gpu.sh r1f'
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 140 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e > $OUT/${TAG}_bench_4gpu.json 2> $OUT/${TAG}_bench_4gpu.err
tail -c 400 $OUT/${TAG}_bench_4gpu.json; tail -3 $OUT/${TAG}_bench_4gpu.err
