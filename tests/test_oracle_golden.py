"""Pins the CPU oracle (oracle/dml_oracle.py) against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import dml_oracle as O

METRIC_CASES = ["kat1", "kat2", "kat3", "kat4", "kat5", "plateau", "allties", "onepos", "oneneg",
                "widerange", "recallsteps"]

# SURVEY.md section 8c known-answer table (generated from the reference during the survey)
SURVEY_KAT = {
    "kat1": (0.7833333333333333, 0.7392857142857143, 0.5),
    "kat2": (0.6834803722976756, 0.5022304006510289, 0.8097560975609757),
    "kat3": (0.8511632974119763, 0.2355103586403926, 0.5661916978332767),
    "kat4": (1.0, 1.0, 0.0),
    "kat5": (0.6666666666666667, 0.6666666666666666, 0.3333333333333333),
}


@pytest.mark.parametrize("name", METRIC_CASES)
@pytest.mark.parametrize("use_sklearn", [False, True])
def test_get_measures(golden, name, use_sklearn):
    g = golden("metrics_kat.npz")
    res = O.get_measures(g[f"{name}_pos"], g[f"{name}_neg"], use_sklearn=use_sklearn)
    np.testing.assert_allclose(np.float64(res), g[f"{name}_res"], rtol=0, atol=1e-12)
    if name in SURVEY_KAT:
        np.testing.assert_allclose(np.float64(res), SURVEY_KAT[name], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", METRIC_CASES)
def test_fpr_at_recall_090(golden, name):
    g = golden("metrics_kat.npz")
    pos, neg = g[f"{name}_pos"], g[f"{name}_neg"]
    labels = np.r_[np.ones(len(pos), np.int32), np.zeros(len(neg), np.int32)]
    got = O.fpr_and_fdr_at_recall(labels, np.r_[pos, neg], 0.90)
    assert got == pytest.approx(float(g[f"{name}_fpr90"]), abs=1e-15)


def test_fpr_rejects_non_binary():
    with pytest.raises(ValueError):
        O.fpr_and_fdr_at_recall(np.array([0, 1, 2]), np.float32([.1, .2, .3]))


def test_eval_ood_measure_wrappers(golden):
    g = golden("metrics_kat.npz")
    conf, seg = g["img_conf"], g["img_seg"]
    np.testing.assert_allclose(O.eval_ood_measure(conf, seg, (13,)), g["img_res_script"], atol=1e-12)
    np.testing.assert_allclose(O.eval_ood_measure_anom_utils(conf, seg, 13), g["img_res_anom_utils"], atol=1e-12)
    np.testing.assert_allclose(O.eval_ood_measure(conf, seg, (13, 5)), g["img_res_script_two_labels"], atol=1e-12)
    m = g["img_mask"]
    np.testing.assert_allclose(O.eval_ood_measure(conf[m], seg, (13,), mask=m), g["img_res_script_masked"], atol=1e-12)
    assert O.eval_ood_measure(conf, np.zeros_like(seg), (13,)) is None


def test_normalization_and_coefficient(golden):
    g = golden("metrics_kat.npz")
    out = O.normalization(g["norm_in"])
    assert out.dtype == np.float32
    np.testing.assert_array_equal(out, g["norm_out"])
    c = O.coefficient_map(g["coef_in"], 0.2)
    assert c.dtype == g["coef_out"].dtype
    np.testing.assert_array_equal(c, g["coef_out"])
    g2 = golden("validate_deeplab.npz")
    np.testing.assert_array_equal(O.normalization(g2["norm_in"]), g2["norm_out"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_head_anomaly(golden, tag):
    g = golden("head_anomaly.npz")
    x = torch.from_numpy(g[f"{tag}_x_low"])
    c = O.make_centers(13)
    np.testing.assert_array_equal(c.numpy(), g[f"{tag}_centers"])
    z_low = O.distance_logits(x, c)
    np.testing.assert_array_equal(z_low.numpy(), g[f"{tag}_z_low"])
    z_up, f_up = O.ppm_head_eval(x, c, (40, 56))
    np.testing.assert_array_equal(z_up.numpy(), g[f"{tag}_z_up"])
    np.testing.assert_array_equal(f_up.numpy(), g[f"{tag}_f_up"])
    # the fp64 truth brackets the fp32 result
    z64 = O.distance_logits_f64(x, c)
    np.testing.assert_allclose(z_low.numpy(), z64.numpy(), rtol=2e-6)


@pytest.mark.parametrize("k", [16, 17, 19])
def test_head_deeplab(golden, k):
    g = golden("head_deeplab.npz")
    feats = torch.from_numpy(g[f"k{k}_features_nhwc"])
    x = feats.permute(0, 3, 1, 2).contiguous()
    c = O.make_centers(k)
    np.testing.assert_array_equal(c.numpy(), g[f"k{k}_centers"])
    np.testing.assert_array_equal(O.distance_logits(x, c).numpy(), g[f"k{k}_logits"])
    np.testing.assert_array_equal(O.features_nhwc(x).numpy(), g[f"k{k}_features_nhwc"])


def test_evaluate_anomaly_end_to_end(golden):
    """Replays evaluate() (multi-scale head -> argmax -> dissum -> per-image metrics -> acc/IoU)
    from the captured stride-8 embeddings and checks every captured reference value."""
    g = golden("evaluate_anomaly.npz")
    c = O.make_centers(13)
    aurocs = []
    for i in range(2):
        lows = [torch.from_numpy(g[f"img{i}_low{s}"]) for s in range(5)]
        seg = g[f"img{i}_seg"]
        scores, _ = O.multiscale_scores(lows, c, seg.shape)
        pred = O.argmax_label(scores)[0]
        np.testing.assert_array_equal(pred, g[f"img{i}_pred"])
        conf = O.score_dissum(scores, 400.0)
        np.testing.assert_array_equal(conf, g[f"img{i}_conf"])
        res = O.eval_ood_measure(conf, seg, (13,))
        np.testing.assert_allclose(res, g[f"img{i}_res"], atol=1e-12)
        aurocs.append(res[0])
        acc, pix = O.accuracy(pred, seg)
        np.testing.assert_allclose([acc, pix], g[f"img{i}_acc"], atol=1e-15)
        inter, union = O.intersection_and_union(pred, seg, 13)
        np.testing.assert_array_equal(inter, g[f"img{i}_inter"])
        np.testing.assert_array_equal(union, g[f"img{i}_union"])
    line = [s for s in g["summary"].tolist() if "mean auroc" in s][0]
    assert float(line.split("mean auroc =")[1].split()[0]) == pytest.approx(np.mean(aurocs), abs=1e-12)


@pytest.mark.parametrize("mode,excl", [("msp", False), ("maxlogit", False), ("dissum", True), ("msp", True), ("maxlogit", True)])
def test_evaluate_anomaly_other_score_modes(golden, mode, excl):
    """`--ood msp` / `--ood maxlogit` / `OOD.exclude_back` of the reference's evaluate()
    (anomaly/eval_ood_traditional.py:212-214,276-278,288-290,302-305), captured from the unmodified reference by
    make_golden.py:gen_evaluate_anomaly_modes: pred, the conf map handed to eval_ood_measure and its result."""
    g = golden("evaluate_anomaly_modes.npz")
    tag = f"{mode}_{'noback' if excl else 'all'}"
    lows = [torch.from_numpy(g[f"{tag}_low{s}"]) for s in range(5)]
    seg = g[f"{tag}_seg"]
    scores, _ = O.multiscale_scores(lows, O.make_centers(13), seg.shape)
    np.testing.assert_array_equal(O.argmax_label(scores)[0], g[f"{tag}_pred"])       # exclude_back never changes pred
    conf = {"msp": O.score_msp, "maxlogit": O.score_maxlogit}[mode](scores, excl) if mode != "dissum" \
        else O.score_dissum(scores, 400.0, excl)
    np.testing.assert_array_equal(conf, g[f"{tag}_conf"])
    np.testing.assert_allclose(O.eval_ood_measure(conf, seg, (13,)), g[f"{tag}_res"], atol=1e-12)


def test_config0_full_shape_streethazards(golden):
    """BASELINE.json configs[0] at its real shape (720x1280, 5 scales, PSPNet-ResNet50dilated random init, run through the
    unmodified reference on the CPU by make_golden.py:gen_config0_full_shape): the oracle replays image 0 from the captured
    stride-8 embeddings -- pred exact, conf exact (subsampled map + float64 sum), per-image AUROC / AUPR / FPR95, acc / IoU."""
    g = golden("config0_full_shape.npz")
    lows = [torch.from_numpy(g[f"img0_low{s}"]) for s in range(5)]
    seg = g["img0_seg"].astype(np.int64)
    assert seg.shape == (720, 1280)
    scores, _ = O.multiscale_scores(lows, O.make_centers(13), seg.shape)
    pred = O.argmax_label(scores)[0]
    np.testing.assert_array_equal(pred, g["img0_pred"])
    np.testing.assert_array_equal(np.bincount(pred.reshape(-1), minlength=13), g["img0_pred_hist"])
    conf = O.score_dissum(scores, 400.0)
    np.testing.assert_array_equal(conf[::8, ::8], g["img0_conf_sub8"])
    assert float(conf.astype(np.float64).sum()) == float(g["img0_conf_sum"])
    assert (conf == 1.0).mean() > 0.001                                    # the 400-clamp plateau is populated (ties)
    np.testing.assert_allclose(O.eval_ood_measure(conf, seg, (13,)), g["img0_res"], atol=1e-12)
    acc, pix = O.accuracy(pred, seg)
    np.testing.assert_allclose([acc, pix], g["img0_acc"], atol=1e-15)
    inter, union = O.intersection_and_union(pred, seg, 13)
    np.testing.assert_array_equal(inter, g["img0_inter"])
    np.testing.assert_array_equal(union, g["img0_union"])
    # the run's summary line is the mean over the 4 images of the per-image results (eval_ood_traditional.py:569,641)
    line = [s for s in g["summary"].tolist() if "mean auroc" in s][0]
    assert float(line.split("mean auroc =")[1].split()[0]) == pytest.approx(np.mean([g[f"img{i}_res"][0] for i in range(4)]), abs=1e-12)


def test_validate_deeplab_npm(golden):
    g = golden("validate_deeplab.npz")
    proto = O.novel_prototype(g["prototypes"].tolist())
    hist = np.zeros((19, 19))
    n_novel = 0
    for i in range(3):
        feats = torch.from_numpy(g[f"img{i}_features_nhwc"])
        x = feats.permute(0, 3, 1, 2).contiguous()
        logits = O.distance_logits(x, O.make_centers(16))
        np.testing.assert_array_equal(logits.numpy(), g[f"img{i}_logits"])
        preds, _, _, _ = O.deeplab_scores(logits, 1000.0)
        preds, _ = O.npm_override(preds, logits, feats, proto, 16, -1.5)
        np.testing.assert_array_equal(preds, g[f"img{i}_preds"])
        n_novel += int((preds == 16).sum())
        targets = O.remap_labels_cityscapes(torch.from_numpy(g[f"img{i}_labels_in"])).numpy()
        np.testing.assert_array_equal(targets, g[f"img{i}_targets"])
        for lt, lp in zip(targets, preds):
            hist += O.fast_hist(lt.flatten(), lp.flatten())
    assert n_novel > 100
    np.testing.assert_array_equal(hist, g["confusion"])
    res = O.seg_results(hist)
    for key in ("Overall Acc", "Mean Acc", "FreqW Acc", "Mean IoU"):
        assert res[key] == pytest.approx(float(g["res_" + key.replace(" ", "_")]), abs=1e-15)
    np.testing.assert_allclose([res["Class IoU"][k] for k in range(19)], g["res_class_iou"], atol=1e-15, equal_nan=True)


def test_plm_merge(golden):
    g = golden("plm.npz")
    outs = []
    for i, k in enumerate((16, 17)):
        feats = torch.from_numpy(g[f"head{i}_features_nhwc"])
        x = feats.permute(0, 3, 1, 2).contiguous()
        z = O.distance_logits(x, O.make_centers(k))
        np.testing.assert_array_equal(z.numpy(), g[f"head{i}_logits"])
        outs.append(z)
    np.testing.assert_array_equal(O.plm_merge(outs, novel_cls=1).numpy(), g["merged_preds"])


def test_loss(golden):
    g = golden("loss.npz")
    emb = torch.from_numpy(g["emb"]).requires_grad_(True)
    tgt = torch.from_numpy(g["target"])
    z = O.distance_logits(emb, O.make_centers(16))
    np.testing.assert_array_equal(z.detach().numpy(), g["logits"])
    loss = O.dml_loss(z, tgt, shipped_early_return=True)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g["shipped_loss"], rtol=1e-6)
    np.testing.assert_allclose(emb.grad.numpy(), g["shipped_grad_emb"], rtol=1e-4, atol=1e-7)
    for tag in ("abg", "vl", "center"):
        a, b, gam = g[f"full_{tag}_coef"]
        emb = torch.from_numpy(g["emb"]).requires_grad_(True)
        z = O.distance_logits(emb, O.make_centers(16))
        feats = emb.permute(0, 2, 3, 1).contiguous()
        loss = O.dml_loss(z, tgt, feats, alpha=a, beta=b, gamma=gam, ignore_index=255)
        loss.backward()
        np.testing.assert_allclose(loss.item(), g[f"full_{tag}_loss"], rtol=2e-6)
        np.testing.assert_allclose(emb.grad.numpy(), g[f"full_{tag}_grad_emb"], rtol=1e-4, atol=1e-7)
    # anomaly train branch: CE(ignore -1) + 0.01 VL
    emb = torch.from_numpy(g["anom_emb"]).requires_grad_(True)
    tgt = torch.from_numpy(g["anom_target"])
    z = O.distance_logits(emb, O.make_centers(13))
    loss = O.dml_loss(z, tgt, alpha=0.01, ignore_index=-1)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g["anom_loss"], rtol=2e-6)
    np.testing.assert_allclose(emb.grad.numpy(), g["anom_grad_emb"], rtol=1e-4, atol=1e-7)


def test_segmetrics(golden):
    g = golden("segmetrics.npz")
    hist = np.zeros((19, 19))
    for gt, pr in zip(g["dl_gt"], g["dl_pred"]):
        for lt, lp in zip(gt, pr):
            hist += O.fast_hist(lt.flatten(), lp.flatten())
    np.testing.assert_array_equal(hist, g["dl_confusion"])
    res = O.seg_results(hist)
    for key in ("Overall Acc", "Mean Acc", "FreqW Acc", "Mean IoU"):
        assert res[key] == pytest.approx(float(g["dl_" + key.replace(" ", "_")]), abs=1e-15)
    np.testing.assert_allclose([res["Class IoU"][k] for k in range(19)], g["dl_class_iou"], atol=1e-15, equal_nan=True)
    acc, pix = O.accuracy(g["an_pred"], g["an_gt"])
    np.testing.assert_allclose([acc, pix], g["an_acc"], atol=1e-15)
    inter, union = O.intersection_and_union(g["an_pred"], g["an_gt"], 13)
    np.testing.assert_array_equal(inter, g["an_inter"])
    np.testing.assert_array_equal(union, g["an_union"])


def test_masked_class_mean_matches_per_class_sums():
    rng = np.random.default_rng(3)
    f = rng.standard_normal((24, 32, 16)).astype(np.float32)
    lab = rng.integers(0, 19, (24, 32))
    lab[:8, :] = 15
    m = O.masked_class_mean(f, lab, 15)
    sums, cnt = O.per_class_sums(f.reshape(-1, 16), lab.reshape(-1), 19)
    np.testing.assert_allclose(m, sums[15] / cnt[15], rtol=1e-5)
    assert O.masked_class_mean(f, np.zeros_like(lab), 15) is None


ROC_CASES = ["continuous", "quantised", "diagonal_pairs", "diagonal_then_negatives", "plateau_first_point", "tiny"]


@pytest.mark.parametrize("name", ROC_CASES)
def test_baseline_roc_measures_golden(golden, name):
    """softmax-baseline evaluator (DeepLabV3Plus-Pytorch/test.py:241-244): the oracle's restatement of
    roc_auc_score / average_precision_score / roc_curve(drop_intermediate) vs scikit-learn's own outputs"""
    g = golden("roc_baseline.npz")
    y, s = g[f"{name}_y"].astype(np.int64), g[f"{name}_s"]
    np.testing.assert_allclose(O.baseline_roc_measures(y, s), g[f"{name}_res"], rtol=0, atol=1e-15)
    assert O.baseline_roc_measures(y, s, 0.90)[2] == float(g[f"{name}_fpr90"])
