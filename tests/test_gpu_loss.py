"""Parity of the fused DCE+VL(+Inter) loss kernels (forward and backward, through the C ABI and the
autograd binding) and of the per-class masked sums against the oracle / reference goldens.  GPU only.
Tolerances: loss 1e-5 relative vs the fp32 oracle (1e-6 vs float64 truth), dx 1e-4 relative
(the reference's own fp32 autograd noise) with an absolute floor, prototype means 1e-6."""
import numpy as np
import pytest
import torch

from oracle import dml_oracle as O

pytestmark = pytest.mark.gpu


def _ref_loss_and_grad(emb, tgt, k, alpha, beta, ignore, dtype=torch.float32, centers=None):
    e = emb.to(dtype).clone().requires_grad_(True)
    c = (O.make_centers(k) if centers is None else centers).to(dtype)
    z = O.distance_logits(e, c)
    loss = O.dml_loss(z, tgt, alpha=alpha, beta=beta, ignore_index=ignore)
    loss.backward()
    return loss.item(), e.grad


@pytest.mark.parametrize("k,ignore,alpha,beta", [(13, -1, 0.01, 0.0), (16, 255, 0.01, 0.01 / 80), (17, 255, 0.0, 0.0), (5, 255, 0.3, 0.02)])
def test_loss_forward_backward_identity(k, ignore, alpha, beta):
    import dml_b200
    g = torch.Generator().manual_seed(k)
    n, h, w = 3, 24, 36
    tgt = torch.randint(0, k, (n, h, w), generator=g)
    emb = torch.randn(n, k, h, w, generator=g) * 0.8 + 3.0 * torch.nn.functional.one_hot(tgt, k).permute(0, 3, 1, 2)
    # some pixels misclassified, some ignored
    flip = torch.rand(n, h, w, generator=g) < 0.2
    tgt = torch.where(flip, torch.randint(0, k, (n, h, w), generator=g), tgt)
    tgt[0, :3] = ignore
    ref32, g32 = _ref_loss_and_grad(emb, tgt, k, alpha, beta, ignore)
    ref64, g64 = _ref_loss_and_grad(emb, tgt, k, alpha, beta, ignore, torch.float64)
    x = emb.cuda().requires_grad_(True)
    loss, parts = dml_b200.dml_loss(x, tgt.cuda(), alpha=alpha, beta=beta, ignore_index=ignore, return_parts=True)
    (loss * 1.7).backward()
    np.testing.assert_allclose(loss.item(), ref32, rtol=1e-5)
    np.testing.assert_allclose(parts[0].item(), ref64, rtol=1e-6)
    assert parts[4].item() == (tgt != ignore).sum().item()
    gx = x.grad.cpu() / 1.7
    scale = g64.abs().max().item()
    np.testing.assert_allclose(gx.numpy(), g64.numpy(), rtol=1e-4, atol=1e-6 * scale)
    np.testing.assert_allclose(gx.numpy(), g32.numpy(), rtol=1e-3, atol=1e-5 * scale)


def test_loss_dense_prototypes():
    import dml_b200
    g = torch.Generator().manual_seed(3)
    k, d, n, h, w = 7, 12, 2, 16, 20
    centers = torch.randn(k, d, generator=g) * 2
    tgt = torch.randint(0, k, (n, h, w), generator=g)
    emb = centers[tgt].permute(0, 3, 1, 2).contiguous() + 0.7 * torch.randn(n, d, h, w, generator=g)
    tgt[1, 4] = 255
    ref64, g64 = _ref_loss_and_grad(emb, tgt, k, 0.05, 0.01, 255, torch.float64, centers)
    x = emb.cuda().requires_grad_(True)
    loss = dml_b200.dml_loss(x, tgt.cuda(), centers=centers.cuda(), alpha=0.05, beta=0.01, ignore_index=255)
    loss.backward()
    np.testing.assert_allclose(loss.item(), ref64, rtol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g64.numpy(), rtol=1e-4, atol=1e-6 * g64.abs().max().item())


def test_loss_golden_reference(golden):
    """values produced by the reference's own loss code (shipped CE/n, intended full form, anomaly branch)"""
    import dml_b200
    g = golden("loss.npz")
    emb, tgt = torch.from_numpy(g["emb"]), torch.from_numpy(g["target"])
    for tag, (a, b) in {"shipped": (0.0, 0.0), "full_abg": tuple(g["full_abg_coef"][:2]), "full_vl": tuple(g["full_vl_coef"][:2])}.items():
        x = emb.cuda().requires_grad_(True)
        loss = dml_b200.dml_loss(x, tgt.cuda(), alpha=float(a), beta=float(b), ignore_index=255)
        loss.backward()
        key = "shipped_loss" if tag == "shipped" else f"{tag}_loss"
        gkey = "shipped_grad_emb" if tag == "shipped" else f"{tag}_grad_emb"
        np.testing.assert_allclose(loss.item(), float(g[key]), rtol=2e-5)
        np.testing.assert_allclose(x.grad.cpu().numpy(), g[gkey], rtol=2e-3, atol=2e-6 * np.abs(g[gkey]).max())
    x = torch.from_numpy(g["anom_emb"]).cuda().requires_grad_(True)
    loss = dml_b200.dml_loss(x, torch.from_numpy(g["anom_target"]).cuda(), alpha=0.01, ignore_index=-1)
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g["anom_loss"]), rtol=2e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["anom_grad_emb"], rtol=2e-3, atol=2e-6 * np.abs(g["anom_grad_emb"]).max())


def test_loss_gradient_finite_differences():
    """central differences in float64 on a small crop (gradcheck-style) pin the analytic backward"""
    import dml_b200
    g = torch.Generator().manual_seed(11)
    k, h, w = 6, 4, 4
    emb = torch.randn(1, k, h, w, generator=g, dtype=torch.float64)
    tgt = torch.randint(0, k, (1, h, w), generator=g)

    def f(e):
        return O.dml_loss(O.distance_logits(e, O.make_centers(k).double()), tgt, alpha=0.1, beta=0.03, ignore_index=255).item()
    num = torch.zeros_like(emb)
    eps = 1e-6
    for i in range(emb.numel()):
        e1, e2 = emb.clone().view(-1), emb.clone().view(-1)
        e1[i] += eps
        e2[i] -= eps
        num.view(-1)[i] = (f(e1.view_as(emb)) - f(e2.view_as(emb))) / (2 * eps)
    x = emb.float().cuda().requires_grad_(True)
    dml_b200.dml_loss(x, tgt.cuda(), alpha=0.1, beta=0.03, ignore_index=255).backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), num.numpy(), rtol=2e-4, atol=1e-6)


def test_loss_is_deterministic():
    import dml_b200
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(4, 13, 96, 128, generator=g)).cuda()
    t = torch.randint(0, 13, (4, 96, 128), generator=g).cuda()
    a = [dml_b200.dml_loss(x, t, alpha=0.01, ignore_index=-1, return_parts=True)[1].clone() for _ in range(3)]
    assert torch.equal(a[0], a[1]) and torch.equal(a[1], a[2])


@pytest.mark.parametrize("nhwc", [False, True])
@pytest.mark.parametrize("d,n_cls", [(16, 19), (13, 14), (5, 3)])
def test_class_sums(nhwc, d, n_cls):
    from dml_b200 import prototypes as P
    g = torch.Generator().manual_seed(d * 10 + n_cls)
    b, h, w = 3, 37, 53
    x = torch.randn(b, d, h, w, generator=g)
    lab = torch.randint(0, n_cls + 2, (b, h, w), generator=g)      # includes out-of-range labels
    lab[0, :10] = 1                                                 # a coherent region
    lab[1] = 255
    xin = x.permute(0, 2, 3, 1).contiguous() if nhwc else x
    sums, counts = P.class_sums(xin.cuda(), lab.cuda(), n_cls, nhwc=nhwc)
    sums, counts = sums.cpu().numpy(), counts.cpu().numpy()
    for i in range(b):
        feats = x[i].permute(1, 2, 0).reshape(-1, d).numpy()
        rs, rc = O.per_class_sums(feats, lab[i].reshape(-1).numpy(), n_cls)
        np.testing.assert_array_equal(counts[i], rc)
        np.testing.assert_allclose(sums[i], rs, rtol=1e-6, atol=1e-4)
    # bit-reproducible
    s2, c2 = P.class_sums(xin.cuda(), lab.cuda(), n_cls, nhwc=nhwc)
    assert np.array_equal(s2.cpu().numpy(), sums)


@pytest.mark.parametrize("d,n_cls,h,w,kind", [(17, 19, 64, 96, "tiles"), (17, 19, 64, 96, "iid"), (16, 32, 48, 64, "rows"),
                                               (13, 40, 40, 52, "iid"), (24, 19, 32, 64, "mixed"), (32, 19, 32, 32, "tiles")])
def test_class_sums_label_patterns(d, n_cls, h, w, kind):
    """the lane-owns-class kernel (n_cls <= 32; 4 pixels per thread when the row length allows it) on coherent tiles,
    independent labels per pixel (its indexed-shuffle path), labels constant along rows, a mix, uint8 and int64
    labels with ignored pixels; n_cls > 32 keeps the shared-memory kernel.  Exact counts, sums to 1e-6."""
    from dml_b200 import prototypes as P
    g = torch.Generator().manual_seed(d + n_cls + h)
    b = 2
    x = torch.randn(b, d, h, w, generator=g) * 2 + 0.5
    if kind == "tiles":
        lab = torch.randint(0, n_cls, (b, h // 16, w // 16), generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2)
    elif kind == "rows":
        lab = torch.randint(0, n_cls, (b, h, 1), generator=g).expand(b, h, w).contiguous()
    elif kind == "iid":
        lab = torch.randint(0, n_cls, (b, h, w), generator=g)
    else:
        lab = torch.randint(0, n_cls, (b, h // 8, w // 8), generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2)
        noise = torch.rand(b, h, w, generator=g) < 0.3
        lab = torch.where(noise, torch.randint(0, n_cls, (b, h, w), generator=g), lab)
    lab = lab.clone()
    lab[torch.rand(b, h, w, generator=g) < 0.1] = 255
    for dtype in (torch.uint8, torch.int64):
        sums, counts = P.class_sums(x.cuda(), lab.to(dtype).cuda(), n_cls)
        sums, counts = sums.cpu().numpy(), counts.cpu().numpy()
        for i in range(b):
            feats = x[i].permute(1, 2, 0).reshape(-1, d).numpy()
            rs, rc = O.per_class_sums(feats, lab[i].reshape(-1).numpy(), n_cls)
            np.testing.assert_array_equal(counts[i], rc)
            np.testing.assert_allclose(sums[i], rs, rtol=1e-6, atol=1e-3)
        s2, _ = P.class_sums(x.cuda(), lab.to(dtype).cuda(), n_cls)
        assert np.array_equal(s2.cpu().numpy(), sums)


def test_novel_prototype_generation_recipe():
    """test_embedding.py:413-419: per support image, mean feature of the novel class if it covers > 5 %"""
    from dml_b200 import prototypes as P
    g = torch.Generator().manual_seed(2)
    feats = torch.randn(4, 24, 32, 16, generator=g)            # NHWC like the head's `features`
    lab = torch.randint(0, 19, (4, 24, 32), generator=g)
    lab[0, :8] = 15
    lab[2, :1, :20] = 15                                         # below 5 %
    lab[3][lab[3] == 15] = 0                                     # absent
    protos = P.novel_prototypes(feats.cuda(), lab.cuda(), cls=15, n_cls=19, min_frac=0.05, nhwc=True)
    ref = [O.masked_class_mean(feats[i].numpy(), lab[i].numpy(), 15) for i in range(4)]
    assert [p is None for p in protos] == [r is None for r in ref]
    for p, r in zip(protos, ref):
        if r is not None:
            np.testing.assert_allclose(p, r, rtol=1e-6, atol=1e-7)
    kept = [r.tolist() for r in ref if r is not None]
    np.testing.assert_allclose(P.prototype_mean(kept), O.novel_prototype(kept), rtol=1e-15)
