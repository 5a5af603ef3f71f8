"""Softmax-baseline OOD evaluator (DeepLabV3Plus-Pytorch/test.py:179-248) on the GPU:
1 - max softmax scores, AUROC / AUPR, and the roc_curve FPR@95 convention (first kept point of the
drop_intermediate ROC curve with tpr >= 0.95) against scikit-learn golden vectors and the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import dml_oracle as O

pytestmark = pytest.mark.gpu

ROC_CASES = ["continuous", "quantised", "diagonal_pairs", "diagonal_then_negatives", "plateau_first_point", "tiny"]


@pytest.mark.parametrize("name", ROC_CASES)
def test_roc_measures_golden(golden, name):
    from dml_b200 import ood
    g = golden("roc_baseline.npz")
    y, s = g[f"{name}_y"], g[f"{name}_s"]
    got = ood.measures_from_scores(torch.from_numpy(s).cuda(), torch.from_numpy(y).cuda(), fpr_convention="roc_curve")
    np.testing.assert_allclose(got, g[f"{name}_res"], rtol=0, atol=1e-12)
    got90 = ood.measures_from_scores(torch.from_numpy(s).cuda(), torch.from_numpy(y).cuda(), recall_level=0.90,
                                     fpr_convention="roc_curve")
    assert got90[2] == pytest.approx(float(g[f"{name}_fpr90"]), abs=1e-15)
    # the other convention is a different number on these cases (anomaly/anom_utils.py:57-65)
    closest = ood.measures_from_scores(torch.from_numpy(s).cuda(), torch.from_numpy(y).cuda())
    ref = O.get_measures(s[y == 1], s[y == 0])
    np.testing.assert_allclose(closest, ref, rtol=0, atol=1e-12)


@pytest.mark.parametrize("n,quant", [(1, 0), (4096, 0), (4097, 16), (100_003, 0), (100_003, 64), (2_000_000, 1000)])
def test_roc_fpr_vs_oracle(n, quant):
    """ragged sizes / ties: the T-th positive may sit in any tile; groups may span tiles"""
    from dml_b200 import ood
    rng = np.random.default_rng(n + quant)
    s = rng.random(n).astype(np.float32)
    if quant:
        s = (np.round(s * quant) / quant).astype(np.float32)
    y = (rng.random(n) < 0.1 + 0.2 * s).astype(np.uint8)
    y[0] = 1
    if n > 1:
        y[1] = 0
    if n == 1:       # a single class: NaN like the batched API (the reference skips such images, test.py:224)
        got = ood.measures_from_scores(torch.from_numpy(s).cuda(), torch.from_numpy(y).cuda(), fpr_convention="roc_curve")
        assert all(np.isnan(v) for v in got)
        return
    got = ood.measures_from_scores(torch.from_numpy(s).cuda(), torch.from_numpy(y).cuda(), fpr_convention="roc_curve")
    np.testing.assert_allclose(got, O.baseline_roc_measures(y, s), rtol=0, atol=1e-12)


def test_batched_segments_roc_fpr():
    """dml_ood_roc_fpr on 6 segments at once (incl. a single-class one) after dml_ood_eval_segments"""
    from dml_b200 import ood
    rng = np.random.default_rng(5)
    n_seg, seg_len = 6, 30_011
    conf = (np.round(rng.random((n_seg, seg_len)) * 200) / 200).astype(np.float32)   # ranked as score = -conf
    y = (rng.random((n_seg, seg_len)) < 0.07).astype(np.uint8)
    y[4] = 0
    ws = ood.OodWorkspace(torch.device("cuda", 0))
    res, stats = ood.eval_segments(torch.from_numpy(conf).cuda(), n_seg, seg_len, positive=torch.from_numpy(y).cuda(),
                                   workspace=ws)
    roc = ood.roc_fpr_after_eval(ws, n_seg, seg_len).cpu().numpy()
    vals, _ = ood.results_to_host(res, stats)
    for i in range(n_seg):
        if i == 4:
            assert np.isnan(roc[i])
        else:
            ref = O.baseline_roc_measures(y[i], -conf[i])
            assert roc[i] == pytest.approx(ref[2], abs=1e-15)
            np.testing.assert_allclose(vals[i, :2], ref[:2], rtol=0, atol=1e-12)


def test_baseline_evaluator_drop_in():
    """softmax_scores + roc_measures == the op sequence of test.py:179-248 replayed on the CPU"""
    import torch.nn.functional as F
    from dml_b200.deeplab import baseline
    g = torch.Generator().manual_seed(2)
    outputs = torch.randn(1, 16, 96, 128, generator=g) * 3
    labels = torch.randint(0, 16, (1, 96, 128), generator=g)
    labels_true = labels.clone()
    labels_true[0, :5] = 255                             # void border: dropped from the evaluation
    labels[0, 20:50, 30:90] = 255                        # held-out classes -> 255 after the dataset remap
    labels[0, :5] = 255
    preds, scores = baseline.softmax_scores(outputs.cuda())
    ref_scores = 1 - F.softmax(outputs, dim=1).max(dim=1)[0].numpy()
    np.testing.assert_allclose(scores.cpu().numpy(), ref_scores, rtol=2e-5, atol=2e-6)
    np.testing.assert_array_equal(preds.cpu().numpy(), outputs.max(dim=1)[1].numpy())
    # exact metric parity on the reference's own score map
    keep = (labels_true != 255).numpy()
    msk = (labels.numpy()[keep] == 255).astype(np.int64)
    ref = O.baseline_roc_measures(msk, ref_scores[keep])
    got = baseline.roc_measures(torch.from_numpy(ref_scores).cuda(), labels.cuda(), labels_true.cuda())
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)
    assert baseline.roc_measures(scores, torch.zeros_like(labels).cuda(), labels_true.cuda()) is None


def test_single_class_input_raises_like_sklearn():
    from dml_b200.deeplab import baseline as B
    s = torch.rand(64, device="cuda")
    lab = torch.full((64,), 255, dtype=torch.int64, device="cuda")
    with pytest.raises(ValueError, match="Only one class"):
        B.roc_measures(s, lab)
    assert B.roc_measures(s, torch.zeros(64, dtype=torch.int64, device="cuda")) is None        # test.py:224: no positive pixel
