#!/usr/bin/env python
"""Times the key exchange of the multi-GPU pooled metric in isolation: all_to_all_single of int32 keys with uneven
splits (every rank keeps 1/G and ships (G-1)/G of KEYS_PER_RANK keys).  Run under torchrun; prints one JSON line
(rank 0).  Used to choose the NCCL p2p channel settings bench.py exports."""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = int(os.environ.get("KEYS_PER_RANK", str(1_382_400_000)))
    src = torch.arange(n, dtype=torch.int32, device=dev)
    dst = torch.empty(n, dtype=torch.int32, device=dev)
    base = n // world
    send = [base] * world
    send[-1] = n - base * (world - 1)
    cnt = torch.tensor(send, dtype=torch.int64, device=dev)
    allc = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt)
    recv = torch.stack(allc)[:, rank].tolist()
    for _ in range(2):
        dist.all_to_all_single(dst[: sum(recv)], src, recv, send)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 5
    for _ in range(reps):
        dist.all_to_all_single(dst[: sum(recv)], src, recv, send)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sent = 4 * (n - send[rank])
        print(json.dumps({"world": world, "keys_per_rank": n, "ms": float(t.item()), "sent_GB_per_rank": sent / 1e9,
                          "GBps_per_rank_per_direction": sent / float(t.item()) / 1e6,
                          "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
