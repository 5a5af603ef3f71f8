"""world_size-2 (and 3) gloo tests on CPU for the multi-GPU pooled metric: splitter choice, uneven all-to-all /
all-gather exchange of sorted shards, carried counts, partial combination.  The kernel-calling steps are replaced by
the NumPy test double (tests/np_ops.py); the result must equal the oracle on the concatenated data bit-for-bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dml_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_shard(rank, case):
    rng = np.random.default_rng(100 + rank)
    n = {0: 5000, 1: 7001, 2: 123}[rank]
    case = case.split("+")[0]
    if case == "empty_rank" and rank == 1:
        n = 0
    conf = rng.random(n).astype(np.float32)
    if case in ("ties", "empty_rank"):
        conf = (np.round(conf * 20) / 20).astype(np.float32)      # the same 21 values on every rank
    conf[rng.random(n) < 0.1] = 1.0                                # clamp plateau shared by all ranks
    gt = rng.integers(0, 14, n).astype(np.int64)
    gt[rng.random(n) < 0.7] = 2
    if case == "skewed":
        conf = (conf * (0.2 + 0.4 * rank)).astype(np.float32)     # shards cover different score ranges
    return conf, gt


def _worker(rank, world, port, case, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dml_b200 import distributed as D
        from tests.np_ops import NumpyOps
        conf, gt = _make_shard(rank, case)
        if "+keys" in case:
            # keys handed over by the per-image evaluation (an ood.KeyPool on the GPU): already generated, and left
            # sorted inside every 1000-pair "image" by the per-image sorts
            ops = NumpyOps()
            keys, stats = ops.make_keys(torch.from_numpy(conf), torch.from_numpy(gt), (13,), 0x80000000)
            k = keys.numpy().view(np.uint32).copy()
            for s0 in range(0, k.size, 1000):
                k[s0:s0 + 1000].sort()
            ex = None
            if "+exchange" in case:
                # what ood.eval_segments(pool=...) does per batch on the GPU: export the batch's positives into the next
                # slot, publish it with the pool's running counts; 8 slots on every rank (short shards pad with empty ones)
                ex = D.PositiveExchange("cpu", slot_keys=1000, max_slots=8)
                ex.begin()
                run = np.zeros(4, np.int64)
                for b in range(8):
                    kb = k[b * 1000:(b + 1) * 1000]
                    pos = (kb[(kb & 1) == 1] >> 1).astype(np.uint32)
                    run[0] += pos.size
                    done = min((b + 1) * 1000, k.size)
                    if kb.size == 0 or ("+hold" in case and rank == 0 and b == 1):
                        ex.publish_empty(torch.from_numpy(run.copy()), done)       # (+hold: a batch that kept its positives)
                        continue
                    slot, cnt = ex.next_slot(1000)
                    slot[:pos.size] = torch.from_numpy(pos.view(np.int32).copy())
                    cnt[0] = pos.size
                    ex.publish(torch.from_numpy(run.copy()), done)
            a, p, f, info = D.pooled_measures(None, None, (13,), mode=mode, ops=ops,
                                              keys_and_stats=(torch.from_numpy(k.view(np.int32)), stats), exchange=ex)
            if ex is not None:
                assert info["positives_from"] == ("allgather" if "+hold" in case else "slots")
        else:
            a, p, f, info = D.pooled_measures(torch.from_numpy(conf), torch.from_numpy(gt), (13,), mode=mode, ops=NumpyOps())
        vals = torch.tensor([[0.5 + 0.1 * rank, 0.2, 0.3], [float("nan")] * 3, [0.7, 0.4, 0.1]], dtype=torch.float64)
        mean = D.mean_of_per_image(vals)
        conf_m = torch.full((3, 3), rank + 1, dtype=torch.int64)
        D.allreduce_counts(conf_m)
        q.put((rank, a, p, f, info, mean, conf_m.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case,mode", [(2, "plain", "alltoall"), (2, "ties", "alltoall"), (2, "ties", "allgather"),
                                             (3, "skewed", "alltoall"), (3, "empty_rank", "allgather"), (2, "empty_rank", "alltoall"),
                                             (2, "plain", "partition"), (2, "ties", "partition"), (3, "skewed", "partition"),
                                             (3, "empty_rank", "partition"), (2, "ties+keys", "partition"),
                                             (3, "empty_rank+keys", "partition"), (2, "plain+keys", "allgather"),
                                             (2, "plain", "rank"), (2, "ties", "rank"), (3, "skewed", "rank"),
                                             (3, "empty_rank", "rank"), (2, "ties+keys", "rank"),
                                             (2, "plain+keys+exchange", "rank"), (3, "empty_rank+keys+exchange", "rank"),
                                             (3, "skewed+keys+exchange+hold", "rank")])
def test_pooled_measures_gloo(world, case, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = [_make_shard(r, case) for r in range(world)]
    conf = np.concatenate([s[0] for s in shards])
    gt = np.concatenate([s[1] for s in shards])
    ref = O.eval_ood_measure(conf, gt, (13,))
    first = sorted(results)[0]
    for r in sorted(results):
        assert (r[1], r[2], r[3]) == (first[1], first[2], first[3]), "ranks must agree bit-for-bit"
        np.testing.assert_allclose([r[1], r[2], r[3]], ref, rtol=0, atol=1e-12)
        assert r[4]["n_pos"] == int((gt == 13).sum()) and r[4]["n_neg"] == int((gt != 13).sum())
        assert r[4]["n_groups"] == (np.unique(conf).size if mode != "rank" else -1)
        # per-image mean across ranks: NaN rows skipped, 2 images per rank counted
        exp = np.mean([[0.5 + 0.1 * k, 0.2, 0.3] for k in range(world)] + [[0.7, 0.4, 0.1]] * world, axis=0)
        np.testing.assert_allclose(r[5][:3], exp, rtol=1e-12)
        assert r[5][3] == 2 * world
        assert r[6] == [[sum(range(1, world + 1))] * 3] * 3
    if mode != "rank":
        assert sum(r[4]["range_keys"] for r in results) == conf.size
    else:
        assert all(r[4]["n_pos_groups"] == np.unique(conf[gt == 13]).size for r in results)


def test_choose_splitters_properties():
    from dml_b200.distributed import choose_splitters
    s = torch.tensor([5, 9, 9, 9, 9, 9, 9, 100, 101, 3000, -1, -1], dtype=torch.int64)
    b = choose_splitters(s, 4)
    assert b[0] == 0 and b[-1] == 1 << 32 and (b[1:] >= b[:-1]).all()
    assert all(int(v) % 2 == 0 for v in b[:-1])          # positive bit cleared: a score never straddles ranges
    assert choose_splitters(torch.tensor([-1, -1], dtype=torch.int64), 2).tolist() == [0, 0, 1 << 32]


def _exchange_contract_worker(port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        from dml_b200 import distributed as D
        ex = D.PositiveExchange("cpu", slot_keys=7, max_slots=2)          # capacity is rounded up to an even number of keys
        out = {"slot_keys": ex.slot_keys, "stride": ex.stride, "errors": []}
        run = torch.zeros(4, dtype=torch.int64)
        for call in (lambda: ex.publish(run, 0),                          # publish without a slot
                     lambda: ex.next_slot(9)):                            # a batch that may export more than a slot holds
            try:
                call()
            except ValueError as e:
                out["errors"].append(str(e))
        keys, cnt = ex.next_slot(4)
        keys[:3] = torch.tensor([5, 1, 9], dtype=torch.int32)
        cnt[0] = 3
        run[0] = 3
        ex.publish(run, 100)
        ex.publish_empty(run, 150)
        try:
            ex.next_slot(1)                                               # more batches than max_slots
        except ValueError as e:
            out["errors"].append(str(e))
        rec = ex.finish()
        hdr = rec[:, :, :D.PositiveExchange.HDR].contiguous().view(torch.int64)
        out["hdr"] = hdr.tolist()
        out["keys"] = rec[0, 0, D.PositiveExchange.HDR: D.PositiveExchange.HDR + 3].tolist()
        ex.begin()
        out["n_after_begin"] = ex.n
        q.put(out)
    finally:
        dist.destroy_process_group()


def test_positive_exchange_contract():
    """slot records, header layout and the error behaviour of distributed.PositiveExchange (one gloo rank)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_exchange_contract_worker, args=(_free_port(), q))
    p.start()
    out = q.get(timeout=120)
    p.join(timeout=60)
    assert p.exitcode == 0
    assert out["slot_keys"] == 8 and out["stride"] == 12 + 8
    assert len(out["errors"]) == 3 and "without next_slot" in out["errors"][0] and "slot capacity" in out["errors"][1] \
        and "max_slots" in out["errors"][2]
    # [slot][rank][(count, n_pos, n_nan, n_oow, -, keys so far)]
    assert out["hdr"] == [[[3, 3, 0, 0, 0, 100]], [[0, 3, 0, 0, 0, 150]]]
    assert out["keys"] == [5, 1, 9] and out["n_after_begin"] == 0
