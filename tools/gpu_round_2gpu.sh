#!/bin/bash
# Two-GPU visit: NCCL tests of the pooled exchange and the N=2 bench line.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 420 -- 'bash tools/gpu_round_2gpu.sh r1e'
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest tests/test_gpu_distributed.py" | tee $OUT/${TAG}_pytest_2gpu.log
timeout 240 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_metrics.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 >> $OUT/${TAG}_pytest_2gpu.log
tail -3 $OUT/${TAG}_pytest_2gpu.log
echo "== bench N=2"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 1 > $OUT/${TAG}_bench_2gpu.json 2> $OUT/${TAG}_bench_2gpu.err
tail -c 400 $OUT/${TAG}_bench_2gpu.json; tail -3 $OUT/${TAG}_bench_2gpu.err
echo "== done"
