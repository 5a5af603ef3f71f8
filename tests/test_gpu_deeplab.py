"""Drop-in surface of the DeepLab sub-project (model wrappers, criterion, StreamSegMetrics, NPM / PLM
evaluation) against values captured from the unmodified reference (tests/golden).  GPU only."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import dml_oracle as O

pytestmark = pytest.mark.gpu


class Preset(nn.Module):
    """stands in for a cuDNN classifier head: returns a preset tensor"""

    def __init__(self, t):
        super().__init__()
        self.t = nn.Parameter(t.clone())

    def forward(self, _):
        return self.t


def test_embedding_model_wrapper(golden):
    from dml_b200.deeplab import _SimpleSegmentationModel_embedding
    g = golden("head_deeplab.npz")
    for k in (16, 17, 19):
        feats = torch.from_numpy(g[f"k{k}_features_nhwc"])
        x = feats.permute(0, 3, 1, 2).contiguous().cuda()
        model = _SimpleSegmentationModel_embedding(nn.Identity(), Preset(x)).cuda().eval()
        with torch.no_grad():
            logits, centers, f = model(torch.zeros(2, 3, *x.shape[-2:], device="cuda"))
        np.testing.assert_allclose(logits.cpu().numpy(), g[f"k{k}_logits"], rtol=1e-5)
        np.testing.assert_array_equal(centers.cpu().numpy(), g[f"k{k}_centers"])
        np.testing.assert_array_equal(f.cpu().numpy(), g[f"k{k}_features_nhwc"])
        assert f.is_contiguous() and tuple(f.shape) == tuple(feats.shape)


def test_embedding_model_trains_through_the_head(golden):
    """autograd through the wrapper: d loss / d embedding equals the reference's autograd (shipped CE/n)"""
    from dml_b200.deeplab import CrossEntropyLoss, _SimpleSegmentationModel_embedding
    g = golden("loss.npz")
    emb = torch.from_numpy(g["emb"]).cuda()
    tgt = torch.from_numpy(g["target"]).cuda()
    head = Preset(emb)
    model = _SimpleSegmentationModel_embedding(nn.Identity(), head).cuda().train()
    logits, _, feats = model(torch.zeros(2, 3, *emb.shape[-2:], device="cuda"))
    loss = CrossEntropyLoss(alpha=0.01, beta=0.01 / 80, gamma=0)(logits, tgt, feats)
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g["shipped_loss"]), rtol=2e-5)
    np.testing.assert_allclose(head.t.grad.cpu().numpy(), g["shipped_grad_emb"], rtol=2e-3,
                               atol=2e-6 * np.abs(g["shipped_grad_emb"]).max())
    np.testing.assert_allclose(logits.detach().cpu().numpy(), g["logits"], rtol=1e-5)


@pytest.mark.parametrize("tag", ["abg", "vl", "center"])
def test_criterion_full_form(golden, tag):
    from dml_b200.deeplab import CrossEntropyLoss
    from dml_b200 import distance_logits
    g = golden("loss.npz")
    a, b, gam = g[f"full_{tag}_coef"]
    emb = torch.from_numpy(g["emb"]).cuda().requires_grad_(True)
    tgt = torch.from_numpy(g["target"]).cuda()
    logits = distance_logits(emb)
    feats = emb.permute(0, 2, 3, 1).contiguous()
    crit = CrossEntropyLoss(alpha=float(a), beta=float(b), gamma=float(gam), shipped_early_return=False)
    loss = crit(logits, tgt, feats)
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g[f"full_{tag}_loss"]), rtol=2e-5)
    ref = g[f"full_{tag}_grad_emb"]
    np.testing.assert_allclose(emb.grad.cpu().numpy(), ref, rtol=2e-3, atol=2e-6 * np.abs(ref).max())


def test_stream_seg_metrics(golden):
    from dml_b200.deeplab import StreamSegMetrics
    g = golden("segmetrics.npz")
    m = StreamSegMetrics(16)
    assert m.n_classes == 19
    m.reset()
    for gt, pr in zip(g["dl_gt"], g["dl_pred"]):
        m.update(gt, pr)                        # numpy in, like the reference callers
    res = m.get_results()
    np.testing.assert_array_equal(m.confusion_matrix, g["dl_confusion"])
    for key in ("Overall Acc", "Mean Acc", "FreqW Acc", "Mean IoU"):
        assert res[key] == pytest.approx(float(g["dl_" + key.replace(" ", "_")]), abs=1e-15)
    np.testing.assert_allclose([res["Class IoU"][k] for k in range(19)], g["dl_class_iou"], atol=1e-15, equal_nan=True)
    assert m.to_str(res) == str(g["dl_to_str"])
    m.reset()
    assert m.get_results()["Class IoU"].keys() == res["Class IoU"].keys()


def test_npm_eval_batch(golden):
    """whole-batch NPM validation step == the reference's batch-1 loop (preds, targets, confusion, scores)"""
    from dml_b200.deeplab import StreamSegMetrics, evaluation as E
    g = golden("validate_deeplab.npz")
    proto = torch.from_numpy(O.novel_prototype(g["prototypes"].tolist()))
    x = torch.cat([torch.from_numpy(g[f"img{i}_features_nhwc"]).permute(0, 3, 1, 2) for i in range(3)]).contiguous().cuda()
    labels = torch.cat([torch.from_numpy(g[f"img{i}_labels_in"]) for i in range(3)]).cuda()
    m = StreamSegMetrics(16)
    m.reset()
    res = E.npm_eval_batch(x, labels, proto, confusion=m.accumulator("cuda"))
    for i in range(3):
        np.testing.assert_array_equal(res["preds"][i].cpu().numpy(), g[f"img{i}_preds"][0])
        np.testing.assert_array_equal(res["targets"][i].cpu().numpy(), g[f"img{i}_targets"][0].astype(np.uint8))
        logits = torch.from_numpy(g[f"img{i}_logits"])
        _, s_soft, _, s_dis = O.deeplab_scores(logits, 1000.0)
        np.testing.assert_allclose(res["scores_auc_softmax"][i].cpu().numpy(), s_soft, rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(res["scores_auc_dis"][i].cpu().numpy(), s_dis, rtol=1e-5, atol=1e-6)
    out = m.get_results()
    np.testing.assert_array_equal(m.confusion_matrix, g["confusion"])
    for key in ("Overall Acc", "Mean Acc", "FreqW Acc", "Mean IoU"):
        assert out[key] == pytest.approx(float(g["res_" + key.replace(" ", "_")]), abs=1e-15)


def test_plm_model_and_merge(golden):
    from dml_b200.deeplab import _SimpleSegmentationModel_embedding_self_distillation, evaluation as E
    p = golden("plm.npz")
    xs = [torch.from_numpy(p[f"head{i}_features_nhwc"]).permute(0, 3, 1, 2).contiguous().cuda() for i in range(2)]
    model = _SimpleSegmentationModel_embedding_self_distillation(nn.Identity(), [Preset(xs[0]), Preset(xs[1])]).cuda().eval()
    assert model.classifier_list == ['classifier', 'classifier_1'] and hasattr(model, 'classifier_1')
    with torch.no_grad():
        logits, centers, feats = model(torch.zeros(2, 3, *xs[0].shape[-2:], device="cuda"))
    for i in range(2):
        np.testing.assert_allclose(logits[i].cpu().numpy(), p[f"head{i}_logits"], rtol=1e-5)
        np.testing.assert_array_equal(centers[i].cpu().numpy(), p[f"head{i}_centers"])
        np.testing.assert_array_equal(feats[i].cpu().numpy(), p[f"head{i}_features_nhwc"])
    merged = E.plm_eval_batch(xs)
    np.testing.assert_array_equal(merged.cpu().numpy(), p["merged_preds"])
    # training pseudo labels vs the oracle restatement of test_self_distillation.py:558-570
    g = torch.Generator().manual_seed(1)
    labels = torch.randint(0, 17, merged.shape, generator=g)
    labels[torch.rand(merged.shape, generator=g) < 0.3] = 255
    outs = [torch.from_numpy(p[f"head{i}_logits"]) for i in range(2)]
    ref = O.plm_pseudo_labels(labels, outs, novel_cls=1)
    got = E.plm_pseudo_labels(labels.cuda(), xs, novel_cls=1)
    np.testing.assert_array_equal(got.cpu().numpy(), ref.numpy())
