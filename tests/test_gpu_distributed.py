"""Pooled metric through the real CUDA ops + NCCL: world_size 1 always, world_size 2 when two GPUs are visible.
Result must equal the single-segment GPU evaluation and the CPU oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dml_oracle as O

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard(rank, n=300_000):
    rng = np.random.default_rng(7 + rank)
    conf = rng.random(n).astype(np.float32)
    conf[:n // 4] = np.round(conf[:n // 4] * 64) / 64
    conf[rng.random(n) < 0.05] = 1.0
    gt = rng.integers(0, 14, n).astype(np.uint8)
    gt[rng.random(n) < 0.6] = 1
    return conf, gt


def _worker(rank, world, port, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dml_b200 import distributed as D
        conf, gt = _shard(rank)
        if "+exchange" in mode:
            # the bench's multi-GPU flow: rank-path batches publish their positives to all ranks as they go
            # (PositiveExchange); "+sortbatch": one batch takes the sort path and keeps its positives -> the pooled stage
            # notices the shortfall in the slot headers and falls back to compaction + all-gather
            from dml_b200 import ood
            pool = ood.KeyPool(conf.size, "cuda", histograms=False)
            pool.exchange = D.PositiveExchange("cuda", slot_keys=6 * 4096, max_slots=4)
            c, g = torch.from_numpy(conf).cuda().view(10, -1), torch.from_numpy(gt).cuda().view(10, -1)
            for rep in range(2):                       # twice: reset() must recycle the slots
                pool.reset()
                for bi, (s0, s1) in enumerate(((0, 6), (6, 10))):
                    m = "sort" if ("+sortbatch" in mode and bi == 1) else "rank"
                    ood.eval_segments(c[s0:s1], s1 - s0, c.shape[1], gt=g[s0:s1], out_labels=(13,), pool=pool, method=m,
                                      pos_capacity=4096)
                a, p, f, info = D.pooled_measures(None, None, (13,), mode="rank", keys_and_stats=(pool.keys, pool.stats[0]),
                                                  exchange=pool.exchange)
            assert info["positives_from"] == ("allgather" if "+sortbatch" in mode else "slots")
        elif mode.endswith("+pool"):
            # the bench's flow: the per-image evaluation (10 "images" of 30 000 pairs, two batches) leaves its keys and
            # counts in an ood.KeyPool; the pooled exchange starts from those keys instead of generating them again
            from dml_b200 import ood
            pool = ood.KeyPool(conf.size, "cuda", histograms=False)
            c, g = torch.from_numpy(conf).cuda().view(10, -1), torch.from_numpy(gt).cuda().view(10, -1)
            for s0, s1 in ((0, 6), (6, 10)):
                ood.eval_segments(c[s0:s1], s1 - s0, c.shape[1], gt=g[s0:s1], out_labels=(13,), pool=pool)
            a, p, f, info = D.pooled_measures(None, None, (13,), mode=mode.split("+")[0],
                                              keys_and_stats=(pool.keys, pool.stats[0]))
        else:
            a, p, f, info = D.pooled_measures(torch.from_numpy(conf).cuda(), torch.from_numpy(gt).cuda(), (13,), mode=mode)
        q.put((rank, a, p, f, info))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
@pytest.mark.parametrize("mode", ["partition", "alltoall", "allgather", "partition+pool", "allgather+pool", "rank", "rank+pool",
                                  "rank+pool+exchange", "rank+pool+exchange+sortbatch"])
def test_pooled_measures_nccl(world, mode):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    shards = [_shard(r) for r in range(world)]
    conf = np.concatenate([s[0] for s in shards])
    gt = np.concatenate([s[1] for s in shards]).astype(np.int64)
    ref = O.eval_ood_measure(conf, gt, (13,))
    for r in results:
        assert (r[1], r[2], r[3]) == (results[0][1], results[0][2], results[0][3])
        np.testing.assert_allclose([r[1], r[2], r[3]], ref, rtol=0, atol=1e-12)
        assert r[4]["n_groups"] == (np.unique(conf).size if not mode.startswith("rank") else -1)


@pytest.mark.parametrize("n,n_buckets", [(0, 3), (1, 2), (4095, 1), (4097, 2), (100_003, 8), (1_000_000, 16), (300_000, 5)])
def test_partition_kernel(n, n_buckets):
    """dml_ood_partition: exact bucket sizes, every bucket holds exactly its key range (as a multiset)"""
    from dml_b200.distributed import CudaOps
    rng = np.random.default_rng(n + n_buckets)
    keys = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    if n > 10:
        keys[: n // 3] = keys[n // 2]                         # a heavy tie
    inner = np.sort(rng.integers(0, 1 << 32, n_buckets - 1, dtype=np.uint64).astype(np.uint32)) & ~np.uint32(1)
    if n_buckets > 3:
        inner[2] = inner[1]                                   # an empty range
    ops = CudaOps(torch.device("cuda", 0))
    out, counts = ops.partition(torch.from_numpy(keys.view(np.int32)).cuda(), torch.from_numpy(inner.view(np.int32).copy()))
    out = out.cpu().numpy().view(np.uint32)
    counts = counts.cpu().numpy()
    b = np.searchsorted(inner, keys, side="right")
    np.testing.assert_array_equal(counts, np.bincount(b, minlength=n_buckets))
    off = 0
    for g in range(n_buckets):
        np.testing.assert_array_equal(np.sort(out[off: off + counts[g]]), np.sort(keys[b == g]))
        off += counts[g]
