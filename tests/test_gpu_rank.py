"""Minority-rank metric path (``dml_ood_rank_segments``: positives sorted, every negative located among them in one
pass; csrc/ood_rank.cu) against the CPU oracle (anomaly/anom_utils.py:25-78 restated) and against the sort path
(``dml_ood_keygen`` + ``dml_ood_eval_segments``).  AUROC / FPR must agree EXACTLY with the sort path (both count in
integers), AUPR to float64 summation order; maps and packed keys must be bit-identical.  GPU only."""
import numpy as np
import pytest
import torch

from oracle import dml_oracle as O

pytestmark = pytest.mark.gpu


def _both(conf, gt, n_seg, seg_len, **kw):
    from dml_b200 import ood
    c = torch.from_numpy(conf).cuda()
    g = torch.from_numpy(gt).cuda()
    r1, s1 = ood.eval_segments(c, n_seg, seg_len, gt=g, out_labels=(13,), method="sort", **kw)
    v1, c1 = ood.results_to_host(r1, s1)
    r2, s2 = ood.eval_segments(c, n_seg, seg_len, gt=g, out_labels=(13,), method="rank", **kw)
    v2, c2 = ood.results_to_host(r2, s2)
    return v1, c1, v2, c2


@pytest.mark.parametrize("n_seg,seg_len,pos_frac", [(1, 1, 0.5), (3, 5, 0.4), (4, 4096, 0.05), (5, 4097, 0.02), (2, 70001, 0.01),
                                                     (7, 12289, 0.2), (1, 300000, 0.01), (150, 2048, 0.03), (3, 65536, 0.3)])
def test_rank_equals_sort_and_oracle(n_seg, seg_len, pos_frac):
    rng = np.random.default_rng(n_seg * 1000 + seg_len)
    conf = rng.random((n_seg, seg_len)).astype(np.float32)
    conf[0] = np.round(conf[0] * 50) / 50                 # quantised: many ties between positives and negatives
    conf[rng.random((n_seg, seg_len)) < 0.1] = 1.0        # clamp plateau
    gt = np.full((n_seg, seg_len), 3, np.int64)
    gt[rng.random((n_seg, seg_len)) < pos_frac] = 13
    if n_seg > 2:
        gt[2] = 3                                         # single-class segment -> NaN row
    if n_seg > 3:
        gt[3] = 13                                        # all positive -> NaN row
    v1, c1, v2, c2 = _both(conf, gt, n_seg, seg_len, pos_capacity=32768)
    for s in range(n_seg):
        ref = O.eval_ood_measure(conf[s], gt[s], (13,))
        if ref is None:
            assert np.isnan(v2[s]).all() and np.isnan(v1[s]).all()
            continue
        np.testing.assert_allclose(v2[s], ref, rtol=0, atol=1e-12)
        assert v2[s, 0] == v1[s, 0] and v2[s, 2] == v1[s, 2]          # integer counting on both paths
        assert abs(v2[s, 1] - v1[s, 1]) <= 1e-13
        assert c2[s, 0] == c1[s, 0] and c2[s, 1] == c1[s, 1] and c2[s, 3] == -1


def test_rank_fused_maps_keys_and_pool_equal_the_sort_path():
    """normalisation, conf / MMSP / mix maps and the packed keys handed to a KeyPool are the sort path's, bit for bit;
    the pooled metric evaluated from a pool filled by the rank path equals the one filled by the sort path."""
    from dml_b200 import ood
    rng = np.random.default_rng(7)
    n_seg, seg_len = 6, 40000
    raw = (rng.random((n_seg, seg_len)).astype(np.float32) * 500).astype(np.float32)
    raw[raw >= 400] = 400
    gt = rng.integers(0, 13, (n_seg, seg_len)).astype(np.uint8)
    gt[rng.random((n_seg, seg_len)) < 0.02] = 13
    mm = np.zeros((n_seg, 4), np.float32)
    mm[:, 0], mm[:, 1] = raw.min(1), raw.max(1)
    msp = (0.2 + 0.8 * rng.random((n_seg, seg_len))).astype(np.float32)
    mm[:, 2], mm[:, 3] = msp.min(1), msp.max(1)
    outs = {}
    for method in ("sort", "rank"):
        conf_out, mmsp_out, mix_out = (torch.empty(n_seg, seg_len, device="cuda") for _ in range(3))
        pool = ood.KeyPool(n_seg * seg_len, "cuda")
        res, stats = ood.eval_segments(torch.from_numpy(raw).cuda(), n_seg, seg_len, gt=torch.from_numpy(gt).cuda(),
                                       out_labels=(13,), minmax=torch.from_numpy(mm).cuda(), minmax_slot=0, conf_out=conf_out,
                                       msp=torch.from_numpy(msp).cuda(), msp_norm_out=mmsp_out, mix_out=mix_out, pool=pool,
                                       method=method)
        vals, _ = ood.results_to_host(res, stats)
        keys = np.sort(pool.keys.cpu().numpy().view(np.uint32).reshape(n_seg, seg_len), axis=1)
        pv, _ = ood.results_to_host(*pool.evaluate())
        outs[method] = (vals, conf_out.cpu().numpy(), mmsp_out.cpu().numpy(), mix_out.cpu().numpy(), keys, pv)
    a, b = outs["sort"], outs["rank"]
    for i in (1, 2, 4):
        np.testing.assert_array_equal(a[i], b[i])
    # the mix map: the rank kernel evaluates the sigmoid coefficient with ex2.approx / rcp.approx (<= 7e-7 absolute on the
    # coefficient, csrc/ood_rank.cuh), the key-generation kernel with expf and an IEEE reciprocal
    np.testing.assert_allclose(b[3], a[3], rtol=0, atol=1e-6)
    conf64, mmsp64 = a[1].astype(np.float64), a[2].astype(np.float64)
    c64 = 1.0 / (1.0 + np.exp(50.0 * (conf64 - 0.2)))
    np.testing.assert_allclose(b[3], c64 * conf64 + (1.0 - c64) * mmsp64, rtol=1e-5, atol=1e-6)      # against float64 directly
    assert np.array_equal(a[0][:, 0], b[0][:, 0]) and np.array_equal(a[0][:, 2], b[0][:, 2])
    np.testing.assert_allclose(a[0][:, 1], b[0][:, 1], atol=1e-13)
    np.testing.assert_array_equal(a[5][:, [0, 2]], b[5][:, [0, 2]])
    np.testing.assert_allclose(a[5][:, 1], b[5][:, 1], atol=1e-13)
    for s in range(n_seg):
        np.testing.assert_allclose(b[0][s], O.eval_ood_measure(O.normalization(raw[s]), gt[s].astype(np.int64), (13,)), atol=1e-12)


def test_rank_multi_pass_and_label_sources():
    """more distinct positive scores than one shared-memory pass holds (12288) -> several passes over the pixels;
    positives given as a mask; int64 labels; an odd segment length (scalar loads)."""
    from dml_b200 import ood
    rng = np.random.default_rng(3)
    seg_len = 200003
    conf = rng.random((2, seg_len)).astype(np.float32)
    pos = rng.random((2, seg_len)) < 0.14                   # ~28 000 positives per segment
    res, stats = ood.eval_segments(torch.from_numpy(conf).cuda(), 2, seg_len, positive=torch.from_numpy(pos).cuda(),
                                   method="rank", pos_capacity=32768)
    vals, counts = ood.results_to_host(res, stats)
    for s in range(2):
        ref = O.eval_ood_measure(conf[s], np.where(pos[s], 13, 0), (13,))
        np.testing.assert_allclose(vals[s], ref, atol=1e-12)
    gt64 = torch.from_numpy(np.where(pos, 13, 2).astype(np.int64)).cuda()
    res2, stats2 = ood.eval_segments(torch.from_numpy(conf).cuda(), 2, seg_len, gt=gt64, out_labels=(13,), method="rank",
                                     pos_capacity=32768)
    v2, _ = ood.results_to_host(res2, stats2)
    np.testing.assert_array_equal(v2, vals)


def test_rank_overflow_is_flagged_and_auto_falls_back():
    from dml_b200 import ood
    rng = np.random.default_rng(9)
    seg_len = 50000
    conf = rng.random((3, seg_len)).astype(np.float32)
    gt = np.full((3, seg_len), 1, np.int64)
    gt[rng.random((3, seg_len)) < 0.01] = 13
    gt[1, rng.random(seg_len) < 0.5] = 13                  # ~25 000 positives > capacity 4096
    c, g = torch.from_numpy(conf).cuda(), torch.from_numpy(gt).cuda()
    res, stats = ood.eval_segments(c, 3, seg_len, gt=g, out_labels=(13,), method="rank", pos_capacity=4096)
    with pytest.raises(ood.MinorityOverflow):
        ood.results_to_host(res, stats)
    st = stats.cpu().numpy()
    assert st[:, 3].tolist() == [0, 1, 0] and st[1, 0] == (gt[1] == 13).sum()
    r = res.cpu().numpy()
    assert np.isnan(r[1, :3]).all() and not np.isnan(r[[0, 2], :3]).any()
    res, stats = ood.eval_segments(c, 3, seg_len, gt=g, out_labels=(13,), method="auto", pos_capacity=4096)
    vals, _ = ood.results_to_host(res, stats)
    for s in range(3):
        np.testing.assert_allclose(vals[s], O.eval_ood_measure(conf[s], gt[s], (13,)), atol=1e-12)


def test_rank_degenerate_groups():
    """all positives tied on one score; all negatives tied; positives below / above every negative; recall levels."""
    from dml_b200 import ood
    cases = []
    conf = np.full(5000, 0.5, np.float32); gt = np.zeros(5000, np.int64); gt[:50] = 13
    cases.append((conf.copy(), gt.copy()))                                              # everything tied
    conf = np.linspace(0, 1, 5000, dtype=np.float32); gt = np.zeros(5000, np.int64); gt[:100] = 13
    cases.append((conf.copy(), gt.copy()))                                              # perfect separation
    gt = np.zeros(5000, np.int64); gt[-100:] = 13
    cases.append((conf.copy(), gt.copy()))                                              # perfectly wrong
    conf = np.where(np.arange(5000) % 2 == 0, 0.25, 0.75).astype(np.float32); gt = np.zeros(5000, np.int64); gt[::7] = 13
    cases.append((conf.copy(), gt.copy()))                                              # two score values only
    conf = np.random.default_rng(1).random(5000).astype(np.float32); gt = np.zeros(5000, np.int64); gt[17] = 13
    cases.append((conf.copy(), gt.copy()))                                              # a single positive
    for level in (0.95, 0.5, 0.999, 1.0):
        for conf, gt in cases:
            res, stats = ood.eval_segments(torch.from_numpy(conf).cuda(), 1, 5000, gt=torch.from_numpy(gt).cuda(),
                                           out_labels=(13,), method="rank", recall_level=level)
            vals, _ = ood.results_to_host(res, stats)
            ref = O.get_measures(-conf[gt == 13], -conf[gt != 13], recall_level=level)
            np.testing.assert_allclose(vals[0], ref, atol=1e-12)


def test_rank_full_size_images_equal_sort_path():
    """the bench's own images (720 x 1280, ~1 % OOD pixels): per-image numbers of the two methods"""
    import bench
    from dml_b200.anomaly.eval_ood import EmbeddingEvaluator
    gen = torch.Generator(device="cuda").manual_seed(21)
    x, gt = bench.synth_chunk_torch(4, 13, 720, 1280, gen, torch.device("cuda"))
    a = EmbeddingEvaluator(num_class=13, out_labels=(13,))(x, gt)
    va, _ = a.host()
    conf_a = a.conf.clone()
    b = EmbeddingEvaluator(num_class=13, out_labels=(13,), method="rank")(x, gt)
    vb, _ = b.host()
    assert torch.equal(conf_a, b.conf)
    assert np.array_equal(va[:, 0], vb[:, 0]) and np.array_equal(va[:, 2], vb[:, 2])
    np.testing.assert_allclose(va[:, 1], vb[:, 1], atol=1e-13)


# ---------------------------------------------------------------------------------------------
# pooled ranking (csrc/ood_pool_rank.cu): positives sorted, negatives bucketed + located, scan over the groups
# ---------------------------------------------------------------------------------------------
def _pooled_both(conf, gt, recall_level=0.95):
    from dml_b200 import ood
    n = conf.size
    c, g = torch.from_numpy(conf.reshape(-1)).cuda(), torch.from_numpy(gt.reshape(-1)).cuda()
    r1, s1 = ood.eval_segments(c, 1, n, gt=g, out_labels=(13,), recall_level=recall_level)
    v1, c1 = ood.results_to_host(r1, s1)
    pool = ood.KeyPool(n, "cuda", histograms=False)
    # fill the pool through the rank path (keys only), then rank the pooled keys
    ood.eval_segments(c, 1, n, gt=g, out_labels=(13,), method="rank", pos_capacity=32768, pool=pool)
    r2, s2 = pool.evaluate(recall_level, method="rank")
    v2, c2 = ood.results_to_host(r2.clone(), s2[:, :3])
    return v1[0], c1[0], v2[0], c2[0]


@pytest.mark.parametrize("n,pos_frac,quant", [(5000, 0.02, 0), (4099, 0.3, 50), (300001, 0.01, 0), (2000000, 0.02, 0),
                                              (3000000, 0.4, 0), (1000003, 0.05, 1000)])
def test_pooled_rank_equals_sort_and_oracle(n, pos_frac, quant):
    rng = np.random.default_rng(n)
    conf = rng.random(n).astype(np.float32)
    if quant:
        conf = (np.round(conf * quant) / quant).astype(np.float32)
    conf[rng.random(n) < 0.07] = 1.0                                   # clamp plateau
    gt = np.full(n, 2, np.int64)
    gt[rng.random(n) < pos_frac] = 13
    v1, c1, v2, c2 = _pooled_both(conf, gt)
    assert v2[0] == v1[0] and v2[2] == v1[2] and abs(v2[1] - v1[1]) <= 1e-13
    assert c2[0] == c1[0] and c2[1] == c1[1]
    if n <= 400000:
        np.testing.assert_allclose(v2, O.eval_ood_measure(conf, gt, (13,)), atol=1e-12)


def test_pooled_rank_degenerate_and_levels():
    from dml_b200 import ood
    n = 100000
    rng = np.random.default_rng(4)
    base = rng.random(n).astype(np.float32)
    cases = []
    gt = np.zeros(n, np.int64); gt[:500] = 13
    cases.append((np.full(n, 0.25, np.float32), gt.copy()))                                  # everything tied
    cases.append((np.sort(base), gt.copy()))                                                 # positives = the lowest conf
    cases.append((np.sort(base)[::-1].copy(), gt.copy()))                                    # positives = the highest conf
    g2 = np.zeros(n, np.int64); g2[777] = 13
    cases.append((base.copy(), g2))                                                          # one positive
    g3 = np.full(n, 13, np.int64); g3[123] = 0
    cases.append((base.copy(), g3))                                                          # one negative
    for level in (0.95, 0.5, 1.0):
        for conf, gt in cases:
            v1, c1, v2, c2 = _pooled_both(conf, gt, level)
            assert v2[0] == v1[0] and v2[2] == v1[2] and abs(v2[1] - v1[1]) <= 1e-13
            np.testing.assert_allclose(v2, O.get_measures(-conf[gt == 13], -conf[gt != 13], recall_level=level), atol=1e-12)
    # single class -> NaN row
    c = torch.from_numpy(base).cuda()
    pool = ood.KeyPool(n, "cuda", histograms=False)
    ood.eval_segments(c, 1, n, gt=torch.zeros(n, dtype=torch.int64, device="cuda"), out_labels=(13,), method="rank", pool=pool)
    r, s = pool.evaluate(method="rank")
    assert np.isnan(r.cpu().numpy()[0, :3]).all()


def test_key_pool_rank_over_batches_equals_sort():
    """per-image batches through the rank path feed one pool; its pooled metric (rank) equals the sort path's"""
    from dml_b200 import ood
    rng = np.random.default_rng(12)
    seg_len, batches = 30000, [3, 2, 4]
    tot = sum(batches) * seg_len
    pools = {m: ood.KeyPool(tot, "cuda") for m in ("sort", "rank")}
    for b in batches:
        conf = rng.random((b, seg_len)).astype(np.float32)
        gt = np.where(rng.random((b, seg_len)) < 0.03, 13, 1).astype(np.uint8)
        for m in ("sort", "rank"):
            ood.eval_segments(torch.from_numpy(conf).cuda(), b, seg_len, gt=torch.from_numpy(gt).cuda(), out_labels=(13,),
                              method=m, pool=pools[m])
    a, _ = ood.results_to_host(*pools["sort"].evaluate())
    r, st = pools["rank"].evaluate(method="rank")
    b_, _ = ood.results_to_host(r, st[:, :3])
    assert a[0, 0] == b_[0, 0] and a[0, 2] == b_[0, 2] and abs(a[0, 1] - b_[0, 1]) <= 1e-13
    # a rank-filled pool evaluated by the sort path (no digit histograms were left: the sort counts them itself)
    c, _ = ood.results_to_host(*pools["rank"].evaluate(method="sort"))
    np.testing.assert_array_equal(a[:, [0, 2]], c[:, [0, 2]])


def test_rank_error_behaviour_and_plain_scores():
    """NaN scores raise like scikit-learn in the reference path; plain scores of either sign (score_kind=1, key window
    chosen from the data) rank like the sort path."""
    from dml_b200 import ood
    rng = np.random.default_rng(2)
    n = 20000
    conf = rng.random(n).astype(np.float32)
    gt = np.where(rng.random(n) < 0.05, 13, 0).astype(np.int64)
    bad = conf.copy()
    bad[77] = np.nan
    res, stats = ood.eval_segments(torch.from_numpy(bad).cuda(), 1, n, gt=torch.from_numpy(gt).cuda(), out_labels=(13,), method="rank")
    with pytest.raises(ValueError, match="NaN"):
        ood.results_to_host(res, stats)
    # maxlogit-like scores: negative, higher = more positive
    score = (-(rng.random(n) * 300 + 50)).astype(np.float32)   # one sign: the packed window spans < 2^31 key values
    score[gt == 13] += 40
    st = torch.empty(4, dtype=torch.int64, device="cuda")
    from dml_b200._lib import check, lib, ptr, stream_ptr
    s = torch.from_numpy(score).cuda()
    check(lib().dml_ood_keystats(ptr(s), 1, 1, n, ptr(st), stream_ptr(s.device)), "dml_ood_keystats")
    kmin = int(st[0].item())
    pos = torch.from_numpy(gt == 13).cuda()
    out = {}
    for m in ("sort", "rank"):
        r, t = ood.eval_segments(s, 1, n, positive=pos, score_kind=1, key_base=kmin, method=m)
        out[m], _ = ood.results_to_host(r, t)
    assert out["rank"][0, 0] == out["sort"][0, 0] and out["rank"][0, 2] == out["sort"][0, 2]
    np.testing.assert_allclose(out["rank"][0], O.get_measures(score[gt == 13], score[gt != 13]), atol=1e-12)
