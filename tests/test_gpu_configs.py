"""BASELINE.json configs 3 and 4 at their FULL sizes (config 2 is bench.py's workload, config 5 its multi-GPU run).
The CPU oracle cannot process 33 M pixels in seconds, so parity is checked (i) exactly against the oracle on
windows cut out of the full-size tensors -- every op on the path is per-pixel, so a window of the output must
equal the oracle's output on the same window of the input -- and (ii) through size-independent properties:
batch == per-image loop, linearity of the counters, determinism, gradient identities.  GPU only."""
import numpy as np
import pytest
import torch

from oracle import dml_oracle as O

pytestmark = pytest.mark.gpu


def _mixture(b, k, h, w, seed, sigma=0.5, novel_mu=None, device="cuda"):
    """x = 3 e_c + sigma N(0,1) on a 64x64 block class map; class id k marks a 'novel' cluster around novel_mu"""
    g = torch.Generator(device=device).manual_seed(seed)
    n_cls = k + (1 if novel_mu is not None else 0)
    cm = torch.randint(0, n_cls, (b, (h + 63) // 64, (w + 63) // 64), generator=g, device=device)
    cm = cm.repeat_interleave(64, 1).repeat_interleave(64, 2)[:, :h, :w].contiguous()
    x = torch.randn(b, k, h, w, generator=g, device=device) * sigma
    base = cm.clamp(max=k - 1)
    x.scatter_add_(1, base.unsqueeze(1), (3.0 * (cm < k).float()).unsqueeze(1))
    if novel_mu is not None:
        x += (cm == k).float().unsqueeze(1) * novel_mu.to(device).float().view(1, k, 1, 1)
    return x, cm


def test_config3_npm_eval_cityscapes_batch16():
    """DeepLab NPM eval, 16 prototypes, Cityscapes shape 1024x2048, batch 16 (33.5 M pixels, 2.1 GB of embeddings)"""
    from dml_b200.deeplab.evaluation import npm_eval_batch, remap_labels
    B, K, H, W = 16, 16, 1024, 2048
    rng = np.random.default_rng(3)
    protos = rng.standard_normal((5, K)) * 1.5                     # 5 stored support prototypes (5-shot)
    mu = torch.from_numpy(O.novel_prototype(protos.tolist()))
    x, cm = _mixture(B, K, H, W, seed=30, sigma=0.2, novel_mu=mu)   # tight clusters: E|noise|^2 = 0.64 < 1.5
    labels = cm.clone()                                            # raw train ids 0..16 (13 = held-out -> 16 after remap)
    labels[:, :8] = 255                                            # ignored border
    labels = labels.to(torch.uint8)
    conf = torch.zeros(19, 19, dtype=torch.int64, device="cuda")
    res = npm_eval_batch(x, labels, mu, confusion=conf)
    torch.cuda.synchronize()
    preds, targets = res["preds"], res["targets"]
    assert preds.shape == (B, H, W) and preds.dtype == torch.uint8
    # (i) windows vs the oracle (reference op order: test_embedding.py:339-350,428-451)
    for (b, y0, x0) in [(0, 0, 0), (5, 500, 1000), (15, 1024 - 48, 2048 - 64), (9, 333, 77)]:
        xw = x[b:b + 1, :, y0:y0 + 48, x0:x0 + 64].cpu().contiguous()
        z = O.distance_logits(xw, O.make_centers(K))
        p = O.argmax_label(z)
        feats = O.features_nhwc(xw)
        p, zn = O.npm_override(p, z, feats, mu.numpy())
        got = preds[b, y0:y0 + 48, x0:x0 + 64].cpu().numpy()
        mism = got != p[0]
        assert mism.mean() < 2e-3                                  # fp64-threshold-adjacent pixels only (SURVEY B: NPM quirk)
        if mism.any():
            zmax = z.max(1)[0][0].double().numpy()
            assert (np.minimum(np.abs(zn[mism] + 1.5), np.abs(zn[mism] - zmax[mism])) < 1e-4).all()
        # raw max-softmax score of the window (1 - max softmax, :340-342)
        np.testing.assert_allclose(res["scores_auc_softmax"][b, y0:y0 + 48, x0:x0 + 64].cpu().numpy(),
                                   1.0 - O.score_msp(z), rtol=0, atol=3e-6)
        np.testing.assert_array_equal(targets[b, y0:y0 + 48, x0:x0 + 64].cpu().numpy(),
                                      O.remap_labels_cityscapes(labels[b:b + 1, y0:y0 + 48, x0:x0 + 64].cpu().long())[0].numpy() & 0xFF)
    # (ii) properties: confusion == bincount of (targets, preds) over valid pixels; batch == per-image loop
    t64, p64 = targets.long(), preds.long()
    valid = t64 < 19
    ref_conf = torch.bincount((t64[valid] * 19 + p64[valid]), minlength=361).view(19, 19)
    assert torch.equal(conf, ref_conf)
    assert int(conf.sum()) == int(valid.sum())
    conf2 = torch.zeros_like(conf)
    for b in (0, 7, 15):
        r1 = npm_eval_batch(x[b:b + 1], labels[b:b + 1], mu, confusion=conf2)
        assert torch.equal(r1["preds"][0], preds[b])
        assert torch.equal(r1["scores_auc_dis"][0], res["scores_auc_dis"][b])      # per-image normalisation
    # EDS complement map is in [0,1] with both ends reached per image
    d = res["scores_auc_dis"].view(B, -1)
    assert float(d.min()) == 0.0 and bool((d.max(1).values == 1.0).all())
    # the planted novel cluster is recovered: most pixels of class id 16 get label 16
    novel_px = cm == K
    assert float((preds[novel_px] == 16).float().mean()) > 0.9


def test_config3_model_wrapper_outputs_full_size():
    """_SimpleSegmentationModel_embedding return convention at 1024x2048: (logits NCHW, centers, features NHWC)"""
    from dml_b200 import dml_head
    B, K, H, W = 2, 16, 1024, 2048
    x, _ = _mixture(B, K, H, W, seed=31)
    out = dml_head(x, want_logits=True, label_dtype=torch.int64, want_features=True)
    assert out.logits.shape == (B, K, H, W) and out.features.shape == (B, H, W, K)
    assert torch.equal(out.features, x.permute(0, 2, 3, 1).contiguous())           # exact copy (network/utils.py:92-93)
    assert torch.equal(out.label, out.logits.argmax(1))
    xw = x[1:2, :, 700:732, 1900:1964].cpu().contiguous()
    z = O.distance_logits(xw, O.make_centers(K))
    np.testing.assert_allclose(out.logits[1:2, :, 700:732, 1900:1964].cpu().numpy(), z.numpy(), rtol=1e-5)


def test_config4_fewshot_plm_loss_prototypes():
    """16+1 5-shot: novel prototype masked mean, PLM merge of the base (16) and novel (17) heads, DCE+VL loss
    forward/backward at crop 768, batch 5"""
    import dml_b200
    from dml_b200 import prototypes
    from dml_b200.deeplab.evaluation import plm_eval_batch
    B, H, W = 5, 768, 768
    x16, cm = _mixture(B, 16, H, W, seed=40)
    x17, _ = _mixture(B, 17, H, W, seed=41)
    labels = cm.clone()
    labels[:, 100:400, 200:600] = 16                               # one large novel region per image (> 5 % of pixels)
    g = torch.Generator(device="cuda").manual_seed(42)
    labels[torch.rand(B, H, W, generator=g, device="cuda") < 0.1] = 255

    # --- (c) masked per-class mean (test_embedding.py:413-419): compare with torch float64 (the kernel adds the
    #     <= 32 fp32 values a warp holds for one class in fp32, then accumulates in float64: mean error << 1e-6)
    sums, counts = prototypes.class_sums(x16, labels.to(torch.uint8), 19)
    for b in (0, 4):
        m = labels[b] == 16
        ref = x16[b].double()[:, m].sum(1)
        np.testing.assert_allclose(sums[b, 16].cpu().numpy() / int(m.sum()), ref.cpu().numpy() / int(m.sum()), rtol=1e-6, atol=1e-8)
        assert int(counts[b, 16]) == int(m.sum())
    protos = prototypes.novel_prototypes(x16.permute(0, 2, 3, 1).contiguous(), labels.to(torch.uint8), 16)
    assert len(protos) == B and all(p is not None for p in protos)
    mw = (labels[0] == 16).cpu().numpy()
    ref0 = O.masked_class_mean(x16[0].permute(1, 2, 0).cpu().numpy(), np.where(mw, 16, 0), 16)
    # the reference recipe is np.mean over a [n,16] fp32 array along axis 0: NumPy accumulates that axis naively in
    # fp32, so over 120 000 rows the REFERENCE carries ~1e-5 relative error; the kernel (float64 accumulation) agrees
    # with the float64 truth to 1e-9 (asserted above) and with the reference to the reference's own accuracy
    np.testing.assert_allclose(np.asarray(protos[0], dtype=np.float64), np.asarray(ref0, dtype=np.float64), rtol=1e-4, atol=1e-6)
    truth0 = x16[0].double()[:, labels[0] == 16].mean(1).cpu().numpy()
    np.testing.assert_allclose(np.asarray(protos[0], dtype=np.float64), truth0, rtol=1e-6, atol=1e-9)

    # --- PLM merge (test_self_distillation.py:292-297) on a window vs the oracle, and batch consistency
    conf = torch.zeros(19, 19, dtype=torch.int64, device="cuda")
    preds = plm_eval_batch([x16, x17], labels.to(torch.uint8), confusion=conf, remap=False)
    for (b, y0, x0) in [(0, 0, 0), (4, 700, 690)]:
        z16 = O.distance_logits(x16[b:b + 1, :, y0:y0 + 60, x0:x0 + 70].cpu().contiguous(), O.make_centers(16))
        z17 = O.distance_logits(x17[b:b + 1, :, y0:y0 + 60, x0:x0 + 70].cpu().contiguous(), O.make_centers(17))
        np.testing.assert_array_equal(preds[b, y0:y0 + 60, x0:x0 + 70].cpu().numpy(), O.plm_merge([z16, z17])[0].numpy())
    t64 = labels.long()
    valid = t64 < 19
    assert torch.equal(conf, torch.bincount(t64[valid] * 19 + preds.long()[valid], minlength=361).view(19, 19))

    # --- (b) DCE + VL + Inter loss, forward/backward on the novel head (alpha 0.01, beta 0.01/80: test_embedding.py:726)
    alpha, beta = 0.01, 0.01 / 80
    tgt = labels.clone()
    xg = x17.clone().requires_grad_(True)
    loss = dml_b200.dml_loss(xg, tgt, alpha=alpha, beta=beta, ignore_index=255)
    loss.backward()
    loss2 = dml_b200.dml_loss(x17, tgt, alpha=alpha, beta=beta, ignore_index=255)
    assert float(loss.detach()) == float(loss2)                            # deterministic reduction
    # oracle on image-aligned windows is not possible for a batch-normalised loss: check it through the identity
    # loss(batch) == the reference formula evaluated with torch float64 ops on the GPU tensors
    z = -((x17.double().unsqueeze(1) - 3.0 * torch.eye(17, device="cuda", dtype=torch.float64).view(1, 17, 17, 1, 1)) ** 2).sum(2)
    validm = tgt != 255
    n_valid = int(validm.sum())
    lse = torch.logsumexp(z, dim=1)
    zy = z.gather(1, tgt.clamp(max=16).unsqueeze(1)).squeeze(1)
    ce = ((lse - zy) * validm).sum() / n_valid
    T = H * W
    vl = ((-zy) * validm).view(B, -1).sum(1).div(T).sum()
    inter = ((z.sum(1) - zy) * validm).view(B, -1).sum(1).div(T).sum()
    ref_loss = (ce + alpha * vl + beta * inter) / B
    np.testing.assert_allclose(float(loss), float(ref_loss), rtol=1e-6)
    # gradient: closed form of SURVEY appendix A.6 in float64
    sm = torch.softmax(z, dim=1)
    onehot = torch.nn.functional.one_hot(tgt.clamp(max=16), 17).permute(0, 3, 1, 2).double()
    gk = ((sm - onehot) / n_valid - (alpha / T) * onehot + (beta / T) * (1 - onehot)) * validm.unsqueeze(1) / B
    dx = -2.0 * (gk.sum(1, keepdim=True) * x17.double() - 3.0 * gk)
    err = (xg.grad.double() - dx).abs().max() / dx.abs().max()
    assert float(err) < 1e-5
    assert bool((xg.grad[:, :, ~validm[0]][0] == 0).all())        # ignored pixels get exactly zero gradient
