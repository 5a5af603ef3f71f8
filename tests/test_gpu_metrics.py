"""Parity of the GPU ranking metrics (sort + tie-aware scan through the C ABI) against the golden
vectors produced by the reference and against the CPU oracle.  GPU only.  Tolerance 1e-6
(north star); observed agreement is ~1e-15 because counts are exact integers."""
import numpy as np
import pytest
import torch

from oracle import dml_oracle as O

pytestmark = pytest.mark.gpu

CASES = ["kat1", "kat2", "kat3", "kat4", "kat5", "plateau", "allties", "onepos", "oneneg", "widerange", "recallsteps"]
TOL = 1e-6


@pytest.mark.parametrize("name", CASES)
def test_get_measures_golden(golden, name):
    from dml_b200.anomaly import anom_utils
    g = golden("metrics_kat.npz")
    res = anom_utils.get_measures(g[f"{name}_pos"], g[f"{name}_neg"])
    np.testing.assert_allclose(np.float64(res), g[f"{name}_res"], rtol=0, atol=TOL)
    # integer-exact counting => far tighter in practice
    np.testing.assert_allclose(np.float64(res), g[f"{name}_res"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_fpr_at_recall_golden(golden, name):
    from dml_b200.anomaly import anom_utils
    g = golden("metrics_kat.npz")
    pos, neg = g[f"{name}_pos"], g[f"{name}_neg"]
    labels = np.r_[np.ones(len(pos), np.int32), np.zeros(len(neg), np.int32)]
    got = anom_utils.fpr_and_fdr_at_recall(labels, np.r_[pos, neg], 0.90)
    assert got == pytest.approx(float(g[f"{name}_fpr90"]), abs=1e-12)


def test_eval_ood_measure_golden(golden):
    from dml_b200.anomaly import anom_utils
    g = golden("metrics_kat.npz")
    conf, seg = g["img_conf"], g["img_seg"]
    np.testing.assert_allclose(anom_utils.eval_ood_measure(conf, seg, 13), g["img_res_anom_utils"], atol=1e-12)
    np.testing.assert_allclose(anom_utils.eval_conf_map(conf, seg, (13, 5)), g["img_res_script_two_labels"], atol=1e-12)
    assert anom_utils.eval_ood_measure(conf, np.zeros_like(seg), 13) is None
    # negative conf (e.g. maxlogit) leaves the non-negative fast path and must still agree
    ref = O.eval_ood_measure(-conf - 1.0, seg, (13,))
    np.testing.assert_allclose(anom_utils.eval_conf_map(-conf - 1.0, seg, (13,)), ref, atol=1e-12)


def test_error_behaviour():
    from dml_b200.anomaly import anom_utils
    with pytest.raises(ValueError):
        anom_utils.fpr_and_fdr_at_recall(np.array([0, 1, 2]), np.float32([.1, .2, .3]))
    with pytest.raises(ValueError):
        anom_utils.get_measures(np.float32([0.1, np.nan]), np.float32([0.3]))


@pytest.mark.parametrize("n_seg,seg_len", [(1, 1), (3, 5), (4, 4096), (5, 4097), (2, 70001), (7, 12289), (1, 300000)])
def test_segments_vs_oracle(n_seg, seg_len):
    """Batched segments (the per-image semantics of eval_ood_traditional.py:566-569) incl. ragged tile tails,
    heavy ties (quantised scores) and a clamp plateau at conf == 1."""
    from dml_b200 import ood
    rng = np.random.default_rng(n_seg * 1000 + seg_len)
    conf = rng.random((n_seg, seg_len)).astype(np.float32)
    conf[0] = np.round(conf[0] * 50) / 50            # quantised: many ties
    conf[rng.random((n_seg, seg_len)) < 0.1] = 1.0      # plateau
    gt = rng.integers(0, 14, (n_seg, seg_len)).astype(np.int64)
    gt[rng.random((n_seg, seg_len)) < 0.8] = 3
    if n_seg > 2:
        gt[2] = 3                                       # single-class segment -> NaN row (None in the reference)
    res, stats = ood.eval_segments(torch.from_numpy(conf).cuda(), n_seg, seg_len, gt=torch.from_numpy(gt).cuda(),
                                   out_labels=(13,))
    vals, counts = ood.results_to_host(res, stats)
    for s in range(n_seg):
        ref = O.eval_ood_measure(conf[s], gt[s], (13,))
        if ref is None:
            assert np.isnan(vals[s]).all()
        else:
            np.testing.assert_allclose(vals[s], ref, rtol=0, atol=1e-12)
            assert counts[s, 0] == (gt[s] == 13).sum() and counts[s, 1] == (gt[s] != 13).sum()
            assert counts[s, 3] == np.unique(conf[s]).size


def test_fused_normalisation_keygen():
    """min-max normalisation fused into key generation reproduces NumPy's fp32 (x-min)/(max-min) bit for bit."""
    from dml_b200 import ood
    rng = np.random.default_rng(5)
    raw = (rng.random((3, 5000)).astype(np.float32) * 500).astype(np.float32)
    raw[raw >= 400] = 400
    gt = rng.integers(0, 14, (3, 5000)).astype(np.uint8)
    mm = np.zeros((3, 4), np.float32)
    mm[:, 0], mm[:, 1] = raw.min(1), raw.max(1)
    msp = (0.2 + 0.8 * rng.random((3, 5000))).astype(np.float32)
    mm[:, 2], mm[:, 3] = msp.min(1), msp.max(1)
    conf_out = torch.empty(3, 5000, device="cuda")
    mmsp_out, mix_out = torch.empty(3, 5000, device="cuda"), torch.empty(3, 5000, device="cuda")
    res, stats = ood.eval_segments(torch.from_numpy(raw).cuda(), 3, 5000, gt=torch.from_numpy(gt).cuda(), out_labels=(13,),
                                   minmax=torch.from_numpy(mm).cuda(), minmax_slot=0, conf_out=conf_out,
                                   msp=torch.from_numpy(msp).cuda(), msp_norm_out=mmsp_out, mix_out=mix_out)
    vals, _ = ood.results_to_host(res, stats)
    for s in range(3):
        conf = O.normalization(raw[s])
        np.testing.assert_array_equal(conf_out[s].cpu().numpy(), conf)
        np.testing.assert_allclose(vals[s], O.eval_ood_measure(conf, gt[s].astype(np.int64), (13,)), atol=1e-12)
        mmsp = O.normalization(msp[s])
        np.testing.assert_array_equal(mmsp_out[s].cpu().numpy(), mmsp)
        np.testing.assert_allclose(mix_out[s].cpu().numpy(), O.score_mix(conf, mmsp), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("batches,seg_len", [([4, 4, 2], 4096), ([1, 3], 4097), ([5], 70001), ([2, 2, 2, 1], 12290), ([3, 1], 5)])
def test_key_pool_equals_separate_pooled_evaluation(batches, seg_len):
    """ood.KeyPool: the pooled metric computed from the keys / digit histograms the per-segment calls leave behind
    is bit-identical to a second, separate evaluation of the concatenated maps as one segment (and equals the
    oracle); the per-segment results are unchanged by the pooling."""
    from dml_b200 import ood
    n_seg = sum(batches)
    rng = np.random.default_rng(n_seg * 977 + seg_len)
    raw = (rng.random((n_seg, seg_len)) * 500).astype(np.float32)
    raw[0] = np.round(raw[0] / 10) * 10                 # heavy ties
    raw[raw >= 400] = 400                               # clamp plateau shared by all segments
    gt = rng.integers(0, 14, (n_seg, seg_len)).astype(np.uint8)
    gt[rng.random((n_seg, seg_len)) < 0.8] = 3
    mm = np.zeros((n_seg, 4), np.float32)
    mm[:, 0], mm[:, 1] = raw.min(1), raw.max(1)
    raw_d, gt_d, mm_d = torch.from_numpy(raw).cuda(), torch.from_numpy(gt).cuda(), torch.from_numpy(mm).cuda()
    conf_d = torch.empty(n_seg, seg_len, device="cuda")
    pool = ood.KeyPool(n_seg * seg_len, "cuda")
    for rep in range(2):                                # second round: reset() reuses the buffers
        pool.reset()
        per_seg, s0 = [], 0
        for nb in batches:
            res, stats = ood.eval_segments(raw_d[s0:s0 + nb], nb, seg_len, gt=gt_d[s0:s0 + nb], out_labels=(13,),
                                           minmax=mm_d[s0:s0 + nb], minmax_slot=0, conf_out=conf_d[s0:s0 + nb], pool=pool)
            per_seg.append(ood.results_to_host(res, stats)[0])
            s0 += nb
        pres, pstats = pool.evaluate()
        pooled, pcounts = ood.results_to_host(pres, pstats)
        sres, sstats = ood.eval_segments(conf_d.view(-1), 1, n_seg * seg_len, gt=gt_d.view(-1), out_labels=(13,))
        separate, scounts = ood.results_to_host(sres, sstats)
        np.testing.assert_array_equal(pooled, separate)
        np.testing.assert_array_equal(pcounts, scounts)
        conf = np.stack([O.normalization(raw[s]) for s in range(n_seg)])
        per_seg = np.concatenate(per_seg)
        for got, c, g in [(pooled[0], conf.reshape(-1), gt.reshape(-1))] + [(per_seg[s], conf[s], gt[s]) for s in range(n_seg)]:
            ref = O.eval_ood_measure(c, g.astype(np.int64), (13,))
            if ref is None:                             # single-class segment
                assert np.isnan(got).all()
            else:
                np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)
    with pytest.raises(ValueError):                     # a partly filled pool cannot be evaluated
        pool.reset()
        pool.evaluate()
    with pytest.raises(ValueError):                     # overflow
        ood.eval_segments(raw_d, n_seg, seg_len, gt=gt_d, out_labels=(13,), pool=ood.KeyPool(seg_len, "cuda"))


def test_large_single_segment_properties():
    """2^24+3 pairs: AUROC(score) + AUROC(-score) == 1 with ties counted half, positives/negatives swap symmetry."""
    from dml_b200 import ood
    n = (1 << 24) + 3
    g = torch.Generator(device="cuda").manual_seed(3)
    score = torch.randint(0, 1 << 16, (n,), device="cuda", generator=g).float() / 65536.0
    pos = (torch.rand(n, device="cuda", generator=g) < 0.05 + 0.1 * score)
    a1, p1, f1 = ood.measures_from_scores(score, pos)
    a2, _, _ = ood.measures_from_scores(-score, pos)
    assert abs(a1 + a2 - 1.0) < 1e-12
    a3, _, _ = ood.measures_from_scores(-score, ~pos)
    assert abs(a3 - a1) < 1e-12
    assert 0.5 < a1 < 1 and 0 < p1 < 1 and 0 < f1 <= 1


def test_long_segment_two_level_scan_matches_oracle():
    """40 M pairs in ONE segment (> 8192 scan tiles: the chunked two-level carry / finalize path of the pooled
    metric) against the CPU oracle, including a run of equal keys that spans several chunks."""
    from dml_b200 import ood
    n = 40_000_003
    g = torch.Generator(device="cuda").manual_seed(11)
    score = torch.rand(n, device="cuda", generator=g)
    score[::7] = (score[::7] * 512).round() / 512          # heavy ties
    score[::13] = 1.0                                       # a plateau spanning many tiles
    pos = torch.rand(n, device="cuda", generator=g) < (0.02 + 0.05 * score)
    a, p, f = ood.measures_from_scores(score, pos)
    ref = O.get_measures(score[pos].cpu().numpy(), score[~pos].cpu().numpy())
    np.testing.assert_allclose([a, p, f], ref, rtol=0, atol=1e-9)
    # a block of equal keys longer than a whole chunk (2048 tiles = 8.4 M keys) must carry across chunk borders
    score2 = score.clone()
    score2[5_000_000:25_000_000] = 0.5
    a2, p2, f2 = ood.measures_from_scores(score2, pos)
    ref2 = O.get_measures(score2[pos].cpu().numpy(), score2[~pos].cpu().numpy())
    np.testing.assert_allclose([a2, p2, f2], ref2, rtol=0, atol=1e-9)
