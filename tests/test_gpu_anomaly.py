"""Drop-in surface of the anomaly sub-project: decoder head module, score block of evaluate(), accuracy /
IoU counters, and the fused batch evaluator, against values captured from the reference.  GPU only."""
import numpy as np
import pytest
import torch

from oracle import dml_oracle as O
from tests.synth import streethazards_like

pytestmark = pytest.mark.gpu


def test_evaluate_replay_matches_reference(golden):
    """replays evaluate() of eval_ood_traditional.py from the captured stride-8 embeddings with the GPU ops"""
    from dml_b200 import distance_logits
    from dml_b200.anomaly import eval_ood, utils
    g = golden("evaluate_anomaly.npz")
    cfg = type("C", (), {"OOD": type("O", (), {"out_labels": (13,)})()})()
    aurocs = []
    for i in range(2):
        seg = g[f"img{i}_seg"]
        scores = torch.zeros(1, 13, *seg.shape, device="cuda")
        for s in range(5):
            x_low = torch.from_numpy(g[f"img{i}_low{s}"]).cuda()
            z = distance_logits(x_low)                                   # CUDA head at stride 8
            scores = scores + torch.nn.functional.interpolate(z, size=seg.shape, mode="bilinear", align_corners=False) / 5
        pred, conf = eval_ood.score_map(scores, "dissum")
        pred = pred[0].cpu().numpy()
        mism = pred != g[f"img{i}_pred"]
        assert mism.mean() < 1e-3                                          # near-ties after bilinear averaging only
        np.testing.assert_allclose(conf[0].cpu().numpy(), g[f"img{i}_conf"], rtol=1e-4, atol=2e-6)
        res = eval_ood.eval_ood_measure(conf[0], seg, cfg)
        np.testing.assert_allclose(res, g[f"img{i}_res"], atol=2e-5)    # scores differ in the last bits -> a few rank swaps
        aurocs.append(res[0])
        # counters on the reference's own prediction: exact
        acc, pix = utils.accuracy(g[f"img{i}_pred"], seg)
        np.testing.assert_allclose([acc, pix], g[f"img{i}_acc"], atol=1e-15)
        inter, union = utils.intersectionAndUnion(g[f"img{i}_pred"], seg, 13)
        np.testing.assert_array_equal(inter, g[f"img{i}_inter"])
        np.testing.assert_array_equal(union, g[f"img{i}_union"])
        # exact metric parity on the reference's own conf map
        np.testing.assert_allclose(eval_ood.eval_ood_measure(g[f"img{i}_conf"], seg, cfg), g[f"img{i}_res"], atol=1e-12)


def test_accuracy_iou_golden(golden):
    from dml_b200.anomaly import utils
    g = golden("segmetrics.npz")
    acc, pix = utils.accuracy(g["an_pred"], g["an_gt"])
    np.testing.assert_allclose([acc, pix], g["an_acc"], atol=1e-15)
    inter, union = utils.intersectionAndUnion(g["an_pred"], g["an_gt"], 13)
    np.testing.assert_array_equal(inter, g["an_inter"])
    np.testing.assert_array_equal(union, g["an_union"])


@pytest.mark.parametrize("mode", ["msp", "maxlogit", "dissum", "mmsp", "mix", "background"])
@pytest.mark.parametrize("exclude_back", [False, True])
def test_score_modes(mode, exclude_back):
    from dml_b200.anomaly import eval_ood
    x, _ = streethazards_like(1, 48, 64, k=13, seed=4)
    scores = O.distance_logits(x, O.make_centers(13))
    pred, conf = eval_ood.score_map(scores.cuda(), mode, exclude_back=exclude_back)
    ref = {"msp": O.score_msp, "maxlogit": O.score_maxlogit, "background": O.score_background,
           "dissum": lambda s, exclude_back: O.score_dissum(s, 400.0, exclude_back),
           "mmsp": O.score_mmsp}
    if mode == "mix":
        want = O.score_mix(O.score_dissum(scores, 400.0, exclude_back), O.score_mmsp(scores, exclude_back))
    else:
        want = ref[mode](scores, exclude_back=exclude_back)
    np.testing.assert_allclose(conf[0].cpu().numpy(), want, rtol=2e-5, atol=3e-6)
    np.testing.assert_array_equal(pred[0].cpu().numpy(), O.argmax_label(scores)[0])


def test_decoder_module_matches_reference_head(golden):
    """PPMDeepsup_embedding drop-in: same state_dict layout, eval and train return conventions"""
    from dml_b200.anomaly.models import PPMDeepsup_embedding
    dec = PPMDeepsup_embedding(num_class=13, fc_dim=32, use_softmax=True).cuda().eval()
    keys = set(dec.state_dict().keys())
    assert {"ppm.0.1.weight", "ppm.3.2.running_var", "conv_last.0.weight", "conv_last.4.bias", "cbr_deepsup.0.weight",
            "conv_last_deepsup.weight"} <= keys
    g = torch.Generator().manual_seed(0)
    conv_out = [torch.randn(1, 16, 9, 12, generator=g).cuda(), torch.randn(1, 32, 9, 12, generator=g).cuda()]
    with torch.no_grad():
        z_up, f_up = dec(conv_out, segSize=(40, 56))
        z_only = dec(conv_out, segSize=(40, 56), output_ft=False)
        emb = dec.conv_last(torch.cat([conv_out[-1]] + [torch.nn.functional.interpolate(p(conv_out[-1]), (9, 12), mode="bilinear", align_corners=False) for p in dec.ppm], 1))
    ref_z, ref_f = O.ppm_head_eval(emb.cpu(), O.make_centers(13), (40, 56))
    np.testing.assert_allclose(z_up.cpu().numpy(), ref_z.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(f_up.cpu().numpy(), ref_f.numpy(), rtol=1e-5, atol=1e-6)
    assert torch.equal(z_only, z_up)
    dec.use_softmax = False
    with torch.no_grad():
        (z_low, deepsup), ft = dec(conv_out)
    assert z_low.shape == (1, 13, 9, 12) and deepsup.shape == (1, 13, 9, 12) and ft.shape[1] == 32 + 4 * 512


def test_embedding_evaluator_batch_equals_per_image_reference():
    """the fused batch pipeline (head -> fused-normalisation key-gen -> segmented sort -> scan) reproduces the
    reference's per-image loop: labels, conf maps, per-image metrics, accuracy / IoU summary"""
    from dml_b200.anomaly.eval_ood import EmbeddingEvaluator, summarize
    x, gt = streethazards_like(4, 96, 160, k=13, seed=8, ignore_rows=3)
    gt[3][gt[3] == 13] = 2                                  # an image without OOD pixels -> skipped like the reference
    ev = EmbeddingEvaluator(num_class=13)
    gt_u8 = gt.clone()
    gt_u8[gt_u8 < 0] = 255
    res = ev(x.cuda(), gt_u8.to(torch.uint8).cuda())
    vals, counts = res.host()
    centers = O.make_centers(13)
    ref_vals, inter_sum, union_sum = [], 0, 0
    for i in range(4):
        z = O.distance_logits(x[i:i + 1], centers)
        pred = O.argmax_label(z)[0]
        conf = O.score_dissum(z, 400.0)
        assert (res.label[i].cpu().numpy() != pred).mean() < 1e-4
        np.testing.assert_allclose(res.conf[i].cpu().numpy(), conf, rtol=1e-5, atol=1e-6)
        r = O.eval_ood_measure(conf, gt[i].numpy(), (13,))
        if r is None:
            assert np.isnan(vals[i]).all()
        else:
            # the metric contract (1e-6; exact counting in practice) holds on IDENTICAL scores: the conf map the kernel wrote
            np.testing.assert_allclose(vals[i], O.eval_ood_measure(res.conf[i].cpu().numpy(), gt[i].numpy(), (13,)), atol=1e-9)
            # against the oracle's own fp32 conf map a last-ulp score difference can swap a positive with a neighbouring
            # negative: one swap moves FPR by 1 / N_neg (6.6e-5 on this 96x160 image)
            np.testing.assert_allclose(vals[i], r, atol=2e-4)
            ref_vals.append(r)
        a, b = O.intersection_and_union(res.label[i].cpu().numpy().astype(np.int64), gt[i].numpy(), 13)
        inter_sum, union_sum = inter_sum + a, union_sum + b
    s = summarize(res.confusion.cpu().numpy(), vals)
    assert s["n_images_scored"] == 3
    np.testing.assert_allclose([s["mean_auroc"], s["mean_aupr"], s["mean_fpr"]], np.mean(ref_vals, axis=0), atol=2e-4)
    ok = ~np.isnan(vals[:, 0])
    np.testing.assert_allclose([s["mean_auroc"], s["mean_aupr"], s["mean_fpr"]], vals[ok, :3].mean(axis=0), atol=1e-12)
    np.testing.assert_allclose(s["iou"], inter_sum / (union_sum + 1e-10), rtol=1e-12)
