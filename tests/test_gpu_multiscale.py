"""Fused multi-scale upsample + average + score head against
(i) the CPU oracle (torch-CPU replay of anomaly/models/models.py:659-661 +
anomaly/eval_ood_traditional.py:192-218,302-305), (ii) values captured from the reference's own
``evaluate()`` and (iii) the same op sequence run with torch's CUDA kernels (the code the
reference executes on a GPU).  GPU only; everything goes through ``dml_multiscale_head_forward``."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import dml_oracle as O

pytestmark = pytest.mark.gpu

SH_SCALES = [(38, 67), (47, 84), (57, 100), (66, 117), (71, 125)]   # 5 StreetHazards scales at stride 8


def _low_embeddings(b, k, sizes, seed, sigma=0.7):
    """stride-8 style embeddings: smooth class layout + noise so that interpolated logits cross"""
    g = torch.Generator().manual_seed(seed)
    outs = []
    for (h, w) in sizes:
        cls = torch.randint(0, k, (b, (h + 5) // 6, (w + 5) // 6), generator=g)
        cls = cls.repeat_interleave(6, 1).repeat_interleave(6, 2)[:, :h, :w]
        x = torch.randn(b, k, h, w, generator=g) * sigma + 3.0 * F.one_hot(cls, k).permute(0, 3, 1, 2).float()
        outs.append(x.contiguous())
    return outs


def _near_tie_fraction(scores_ref, label):
    """fraction of mismatching pixels whose top-2 reference logits are NOT within 1e-5 relative"""
    top2 = torch.topk(scores_ref, 2, dim=1).values
    near = (top2[:, 0] - top2[:, 1]).abs() <= 1e-5 * top2[:, 0].abs()
    mism = torch.from_numpy(label.astype(np.int64)) != scores_ref.argmax(1)
    return float((mism & ~near).float().mean()), float(mism.float().mean())


# `exact`: torch's CPU upsample kernel evaluates h0*(w0*a + w1*b) + h1*(w0*c + w1*d) with ONE fixed FMA contraction
# for outputs of realistic size (checked up to 720x1280 when this test was written); the kernel pins the same
# contraction, so the result is bit-identical.  For tiny outputs (a few thousand pixels per plane) torch takes a
# differently contracted code path and agreement is to ~2 ulp instead.
@pytest.mark.parametrize("k,size,sizes,exact", [
    (13, (96, 160), [(5, 9), (7, 11), (8, 13), (9, 15), (10, 17)], True),
    (13, (90, 121), [(12, 16), (23, 31)], True),                 # odd width -> 1 pixel per thread
    (16, (64, 64), [(64, 64)], True),                            # same size: identity interpolation
    (19, (50, 70), [(7, 9), (60, 90), (3, 4)], False),           # one scale is a DOWN-sampling
    (5, (33, 48), [(4, 6)] * 8, False),                          # maximum number of scales
    (13, (180, 320), [(19, 34), (24, 42), (29, 50), (33, 59), (36, 63)], True),   # StreetHazards scales / 2
])
def test_scores_match_cpu_oracle(k, size, sizes, exact):
    from dml_b200 import dml_head, dml_multiscale_head
    embs = _low_embeddings(2, k, sizes, seed=len(sizes) + k)
    centers = O.make_centers(k)
    ref_scores, ref_ft = O.multiscale_scores(embs, centers, size)
    z_list = [dml_head(e.cuda(), want_logits=True, label_dtype=None).logits for e in embs]
    out = dml_multiscale_head(z_list, size, want_scores=True, label_dtype=torch.int64, want_eds=True, eds_clamp=400.0,
                              want_msp=True, want_maxlogit=True, want_minmax=True)
    got = out.logits.cpu()
    # tolerance of the north star: 1e-5 relative on distances (logits are sums of distances: scale by |z|max per pixel)
    scale = ref_scores.abs().amax(1, keepdim=True)
    assert float(((got - ref_scores).abs() / scale).max()) < 1e-5
    bad, mism = _near_tie_fraction(ref_scores, out.label.cpu().numpy())
    assert bad == 0.0 and mism < 1e-3
    eds_ref = np.minimum(-ref_scores.sum(1).numpy(), 400.0)
    np.testing.assert_allclose(out.eds.cpu().numpy(), eds_ref, rtol=1e-5)
    np.testing.assert_allclose(out.msp.cpu().numpy(), O.score_msp(ref_scores), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(out.maxlogit.cpu().numpy(), O.score_maxlogit(ref_scores), rtol=1e-5, atol=1e-5)
    mm = out.minmax.cpu().numpy()
    np.testing.assert_array_equal(mm[:, 0], out.eds.cpu().numpy().reshape(2, -1).min(1))
    np.testing.assert_array_equal(mm[:, 1], out.eds.cpu().numpy().reshape(2, -1).max(1))
    # ft1 accumulator (:209-210): same kernel fed with the raw embeddings.  The kernel pins torch's FMA
    # contraction and divides with correct rounding, so on identical inputs it is BIT-identical to torch-CPU
    from dml_b200 import multiscale_average
    ft = multiscale_average([e.cuda() for e in embs], size).cpu()
    np.testing.assert_allclose(ft.numpy(), ref_ft.numpy(), rtol=1e-6, atol=1e-6)
    # ... and so are the scores / labels when the stride-8 logits are the oracle's own
    z_ref = [O.distance_logits(e, centers).cuda() for e in embs]
    same = dml_multiscale_head(z_ref, size, want_scores=True, label_dtype=torch.int64)
    np.testing.assert_allclose(same.logits.cpu().numpy(), ref_scores.numpy(), rtol=1e-6)
    if exact:
        assert torch.equal(ft, ref_ft)
        assert torch.equal(same.logits.cpu(), ref_scores)
        np.testing.assert_array_equal(same.label.cpu().numpy(), O.argmax_label(ref_scores))


@pytest.mark.parametrize("recip", [True, False])
def test_against_torch_cuda_replay(recip):
    """the reference's GPU op sequence: F.interpolate (CUDA upsample_bilinear2d) and `/ 5` (CUDA div-by-scalar =
    multiplication by fl(1/5)), accumulated in order.  `reciprocal_average=True` reproduces it to the last bit
    up to FMA contraction choices; the correctly rounded division differs by <= 1 ulp per term."""
    from dml_b200 import dml_multiscale_head
    g = torch.Generator().manual_seed(5)
    z_list = [(-(torch.rand(2, 13, h, w, generator=g) * 40 + 0.1)).cuda() for (h, w) in SH_SCALES]
    size = (360, 640)
    scores = torch.zeros(2, 13, *size, device="cuda")
    for z in z_list:
        scores = scores + F.interpolate(z, size=size, mode="bilinear", align_corners=False) / len(z_list)
    out = dml_multiscale_head(z_list, size, reciprocal_average=recip, want_scores=True, label_dtype=torch.int64)
    diff = (out.logits - scores).abs()
    rel = float((diff / scores.abs()).max())
    exact = float((out.logits == scores).float().mean())
    print(f"reciprocal_average={recip}: max rel diff {rel:.3e}, bit-exact fraction {exact:.6f}")
    assert rel < (2e-6 if recip else 4e-6)
    if recip:
        assert exact > 0.99
    assert float((out.label != scores.argmax(1)).float().mean()) < 1e-4


def test_golden_evaluate_replay_fused(golden):
    """evaluate() of eval_ood_traditional.py replayed from the captured stride-8 embeddings: distances at
    stride 8 (CUDA head), then ONE fused upsample+average+score kernel; conf / pred / metrics vs the reference"""
    from dml_b200 import distance_logits
    from dml_b200.anomaly import eval_ood
    g = golden("evaluate_anomaly.npz")
    cfg = type("C", (), {"OOD": type("O", (), {"out_labels": (13,)})()})()
    for i in range(2):
        seg = g[f"img{i}_seg"]
        z_list = [distance_logits(torch.from_numpy(g[f"img{i}_low{s}"]).cuda()) for s in range(5)]
        pred, conf = eval_ood.multiscale_score_map(z_list, seg.shape, "dissum")
        assert (pred[0].cpu().numpy() != g[f"img{i}_pred"]).mean() < 1e-3
        np.testing.assert_allclose(conf[0].cpu().numpy(), g[f"img{i}_conf"], rtol=1e-4, atol=2e-6)
        res = eval_ood.eval_ood_measure(conf[0], seg, cfg)
        np.testing.assert_allclose(res, g[f"img{i}_res"], atol=2e-5)


@pytest.mark.parametrize("mode", ["msp", "maxlogit", "dissum", "mmsp", "mix", "background"])
@pytest.mark.parametrize("exclude_back", [False, True])
def test_fused_score_modes_equal_unfused(mode, exclude_back):
    """multiscale_score_map == score_map(multiscale_scores(...)) bit for bit (same kernel arithmetic)"""
    from dml_b200 import dml_head
    from dml_b200.anomaly import eval_ood
    embs = _low_embeddings(1, 13, [(9, 12), (11, 15), (14, 19)], seed=11)
    z_list = [dml_head(e.cuda(), want_logits=True, label_dtype=None).logits for e in embs]
    size = (72, 96)
    pred_f, conf_f = eval_ood.multiscale_score_map(z_list, size, mode, exclude_back=exclude_back)
    scores = eval_ood.multiscale_scores(z_list, size)
    pred_u, conf_u = eval_ood.score_map(scores, mode, exclude_back=exclude_back)
    assert torch.equal(pred_f, pred_u)
    assert torch.equal(conf_f, conf_u)
    ref_scores, _ = O.multiscale_scores(embs, O.make_centers(13), size)
    if mode == "dissum":
        np.testing.assert_allclose(conf_f[0].cpu().numpy(), O.score_dissum(ref_scores, 400.0, exclude_back), rtol=2e-5, atol=3e-6)


def test_multiscale_evaluator_pipeline():
    """stride-8 embeddings of 5 scales -> labels, conf, confusion, exact per-image metrics == oracle pipeline"""
    from dml_b200.anomaly.eval_ood import MultiScaleEvaluator
    b, k, size = 3, 13, (144, 256)
    sizes = [(8, 13), (10, 17), (12, 20), (14, 24), (15, 25)]
    embs = _low_embeddings(b, k, sizes, seed=21, sigma=0.5)
    g = torch.Generator().manual_seed(22)
    gt = torch.randint(0, k, (b, size[0] // 16, size[1] // 16), generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2)
    gt[:, 40:70, 100:160] = 13
    gt[:, :2] = 255
    ev = MultiScaleEvaluator(num_class=k)
    res = ev([e.cuda() for e in embs], gt.to(torch.uint8).cuda())
    vals, _ = res.host()
    ref_scores, _ = O.multiscale_scores(embs, O.make_centers(k), size)
    gt_ref = gt.clone().long()
    gt_ref[gt_ref == 255] = -1
    conf_all = 0
    for i in range(b):
        conf = O.score_dissum(ref_scores[i:i + 1], 400.0)
        np.testing.assert_allclose(res.conf[i].cpu().numpy(), conf, rtol=2e-5, atol=2e-6)
        pred = O.argmax_label(ref_scores[i:i + 1])[0]
        assert (res.label[i].cpu().numpy() != pred).mean() < 1e-3
        # exact metric parity on OUR conf map (the scan is exact); close to the reference's on its own map
        np.testing.assert_allclose(vals[i], O.eval_ood_measure(res.conf[i].cpu().numpy(), gt_ref[i].numpy(), (13,)), atol=1e-9)
        np.testing.assert_allclose(vals[i], O.eval_ood_measure(conf, gt_ref[i].numpy(), (13,)), atol=5e-5)
        lab = res.label[i].cpu().numpy().astype(np.int64)
        valid = gt_ref[i].numpy() >= 0
        conf_all = conf_all + np.bincount((gt_ref[i].numpy()[valid] * k + lab[valid]), minlength=(k + 1) * k).reshape(k + 1, k)
    np.testing.assert_array_equal(res.confusion.cpu().numpy(), conf_all)


def test_full_streethazards_shape_properties():
    """BASELINE config-1 shape (720x1280, 5 scales, K=13) at full size: size-independent properties --
    a single scale of the output size is the identity; replicating one scale S times equals that scale;
    batch entries are independent of their neighbours"""
    from dml_b200 import dml_multiscale_head
    g = torch.Generator().manual_seed(9)
    z_list = [(-(torch.rand(2, 13, h, w, generator=g) * 30)).cuda() for (h, w) in SH_SCALES]
    size = (720, 1280)
    out = dml_multiscale_head(z_list, size, want_scores=True, label_dtype=torch.uint8, want_eds=True, eds_clamp=400.0)
    assert out.logits.shape == (2, 13, 720, 1280) and bool(torch.isfinite(out.logits).all())
    assert torch.equal(out.label.long(), out.logits.argmax(1))
    one = dml_multiscale_head([z[1:2] for z in z_list], size, want_scores=True, label_dtype=torch.uint8)
    assert torch.equal(one.logits[0], out.logits[1]) and torch.equal(one.label[0], out.label[1])
    full = out.logits[:1].contiguous()
    ident = dml_multiscale_head([full], size, want_scores=True, label_dtype=None)
    assert torch.equal(ident.logits, full)
    # power-of-two replication: v/4 summed 4 times is exact
    rep = dml_multiscale_head([z_list[0]] * 4, size, want_scores=True, label_dtype=None)
    single = dml_multiscale_head([z_list[0]], size, want_scores=True, label_dtype=None)
    assert torch.equal(rep.logits, single.logits)


def test_argument_errors():
    import dml_b200
    from dml_b200 import dml_multiscale_head
    z = torch.zeros(1, 13, 4, 4, device="cuda")
    with pytest.raises(ValueError):
        dml_multiscale_head([], (8, 8))
    with pytest.raises(ValueError):
        dml_multiscale_head([z] * 9, (8, 8))
    with pytest.raises(ValueError):
        dml_multiscale_head([z, torch.zeros(1, 12, 4, 4, device="cuda")], (8, 8))
    with pytest.raises(dml_b200.DmlError):
        dml_multiscale_head([torch.zeros(1, 13, 4, 4)], (8, 8))
    empty = dml_multiscale_head([z[:0]], (8, 8))
    assert empty.label.shape == (0, 8, 8)
