"""GPU: anomaly.lib.nn.SynchronizedBatchNorm2d (dml_bn_* kernels + NCCL all-reduce) against the reference module's outputs
(golden), the float64 oracle, and -- with two GPUs -- a batch split over two ranks against the same batch on one rank
(anomaly/lib/nn/modules/batchnorm.py:57-139)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import syncbn_oracle as S

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "syncbn.npz")


def _module(C, weight=None, bias=None, **kw):
    from dml_b200.anomaly.lib.nn import SynchronizedBatchNorm2d
    m = SynchronizedBatchNorm2d(C, always_sync=True, **kw).cuda()
    with torch.no_grad():
        if weight is not None:
            m.weight.copy_(torch.as_tensor(weight))
            m.bias.copy_(torch.as_tensor(bias))
    return m


def test_two_training_steps_equal_the_reference_module():
    g = np.load(GOLD)
    C = g["weight"].size
    m = _module(C, g["weight"], g["bias"])
    m.train()
    for step in range(2):
        x = torch.from_numpy(g[f"x{step}"]).cuda().requires_grad_(True)
        m.zero_grad()
        y = m(x)
        (y * torch.from_numpy(g[f"g{step}"]).cuda()).sum().backward()
        np.testing.assert_allclose(y.detach().cpu().numpy(), g[f"y{step}"], rtol=2e-5, atol=2e-5)
        ref_dx = g[f"dx{step}"]
        scale = np.abs(ref_dx).max(axis=(0, 2, 3), keepdims=True) + 1e-30
        assert (np.abs(x.grad.cpu().numpy() - ref_dx) / scale).max() < 2e-5
        np.testing.assert_allclose(m.weight.grad.cpu().numpy(), g[f"dw{step}"], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(m.bias.grad.cpu().numpy(), g[f"db{step}"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(m.running_mean.cpu().numpy(), g[f"running_mean{step}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(m.running_var.cpu().numpy(), g[f"running_var{step}"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(m._tmp_running_mean.cpu().numpy(), g[f"tmp_running_mean{step}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(m._running_iter.cpu().numpy(), g[f"running_iter{step}"], rtol=1e-6)
    m.eval()                                                    # evaluation mode: F.batch_norm on the moving averages, like the reference
    y = m(torch.from_numpy(g["x0"]).cuda())
    np.testing.assert_allclose(y.detach().cpu().numpy(), g["y_eval"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape", [(2, 3, 1, 5), (3, 16, 33, 47), (4, 64, 90, 160), (1, 5, 7, 3), (2, 2048, 12, 20)])
@pytest.mark.parametrize("affine", [True, False])
def test_forward_backward_equal_the_oracle(shape, affine):
    B, C, H, W = shape
    rng = np.random.default_rng(B * 1000 + C)
    x = (rng.standard_normal(shape) * rng.uniform(0.5, 3, (1, C, 1, 1)) + rng.uniform(-2, 2, (1, C, 1, 1))).astype(np.float32)
    if C > 2:
        x[:, 2] = 0.25                                           # clamp channel
    gy = rng.standard_normal(shape).astype(np.float32)
    w = rng.uniform(0.5, 1.5, C).astype(np.float32) if affine else None
    b = rng.standard_normal(C).astype(np.float32) if affine else None
    m = _module(C, w, b, affine=affine)
    m.train()
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    y = m(xt)
    (y * torch.from_numpy(gy).cuda()).sum().backward()
    st = S.SyncBNState(C)
    y_ref, cache = S.forward(x, w, b, 1e-5, st)
    dx_ref, dw_ref, db_ref = S.backward(gy, cache)
    np.testing.assert_allclose(y.detach().cpu().numpy(), y_ref, rtol=1e-5, atol=1e-5)
    scale = np.abs(dx_ref).max(axis=(0, 2, 3), keepdims=True) + 1e-30
    assert (np.abs(xt.grad.cpu().numpy() - dx_ref) / scale).max() < 1e-5
    if affine:
        np.testing.assert_allclose(m.weight.grad.cpu().numpy(), dw_ref, rtol=1e-4, atol=1e-4 * np.abs(dw_ref).max())
        np.testing.assert_allclose(m.bias.grad.cpu().numpy(), db_ref, rtol=1e-4, atol=1e-4 * np.abs(db_ref).max())
    np.testing.assert_allclose(m.running_mean.cpu().numpy(), st.running_mean, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(m.running_var.cpu().numpy(), st.running_var, rtol=1e-4, atol=1e-6)


def test_non_parallel_and_errors():
    from dml_b200 import DmlError
    from dml_b200.anomaly.lib.nn import SynchronizedBatchNorm2d, convert_model, patch_replication_callback
    m = SynchronizedBatchNorm2d(4).cuda().train()               # one rank, no always_sync: PyTorch's implementation
    x = torch.randn(2, 4, 3, 3, device="cuda")
    ref = torch.nn.BatchNorm2d(4, momentum=0.001).cuda().train()
    torch.testing.assert_close(m(x).detach(), ref(x).detach())
    with pytest.raises(ValueError):
        SynchronizedBatchNorm2d(4, always_sync=True).cuda().train()(torch.randn(2, 4, 3, device="cuda"))
    with pytest.raises(DmlError):
        SynchronizedBatchNorm2d(4, always_sync=True).train()(torch.randn(2, 4, 3, 3))          # CPU tensor: no fallback
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.BatchNorm2d(4), torch.nn.ReLU()).cuda()
    net = convert_model(net)
    assert isinstance(net[1], SynchronizedBatchNorm2d)
    assert patch_replication_callback(net) is net


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch():
    rng = np.random.default_rng(77)
    x = (rng.standard_normal((6, 8, 21, 34)) * 2 + 1).astype(np.float32)
    gy = rng.standard_normal(x.shape).astype(np.float32)
    return x, gy, rng.uniform(0.5, 1.5, 8).astype(np.float32), rng.standard_normal(8).astype(np.float32)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dml_b200.anomaly.lib.nn import SynchronizedBatchNorm2d
        x, gy, w, b = _batch()
        lo, hi = (0, 4) if rank == 0 else (4, 6)                # uneven shards: the element count is all-reduced too
        m = SynchronizedBatchNorm2d(8).cuda().train()
        with torch.no_grad():
            m.weight.copy_(torch.from_numpy(w))
            m.bias.copy_(torch.from_numpy(b))
        xt = torch.from_numpy(x[lo:hi]).cuda().requires_grad_(True)
        y = m(xt)
        (y * torch.from_numpy(gy[lo:hi]).cuda()).sum().backward()
        q.put((rank, y.detach().cpu().numpy(), xt.grad.cpu().numpy(), m.weight.grad.cpu().numpy(), m.bias.grad.cpu().numpy(),
               m.running_mean.cpu().numpy(), m.running_var.cpu().numpy()))
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_the_global_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x, gy, w, b = _batch()
    st = S.SyncBNState(8)
    y_ref, cache = S.forward(x, w, b, 1e-5, st)
    dx_ref, dw_ref, db_ref = S.backward(gy, cache)
    y = np.concatenate([res[0][1], res[1][1]])
    dx = np.concatenate([res[0][2], res[1][2]])
    np.testing.assert_allclose(y, y_ref, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dx, dx_ref, rtol=1e-4, atol=1e-5 * np.abs(dx_ref).max())
    np.testing.assert_allclose(res[0][3] + res[1][3], dw_ref, rtol=1e-4, atol=1e-4)      # the ranks' own sums add up (DDP's job)
    np.testing.assert_allclose(res[0][4] + res[1][4], db_ref, rtol=1e-4, atol=1e-4)
    for r in res:
        np.testing.assert_allclose(r[5], st.running_mean, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(r[6], st.running_var, rtol=1e-4, atol=1e-6)
