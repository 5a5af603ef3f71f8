#!/usr/bin/env python
"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE
(/root/reference, read-only) on small seeded inputs.

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

Outputs (committed): metrics_kat.npz, head_anomaly.npz, head_deeplab.npz,
evaluate_anomaly.npz, evaluate_anomaly_modes.npz, config0_full_shape.npz, validate_deeplab.npz, loss.npz, segmetrics.npz, roc_baseline.npz.
Every array in them was produced by reference code, never by the oracle or the
product.  Versions at generation time are stored in ``meta.json``.
"""
from __future__ import annotations

import inspect
import io
import json
import os
import sys
import textwrap
import contextlib
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_loader import reference, CfgNode  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrays.items()})


# --------------------------------------------------------------------------- #
def metric_cases():
    """Seeded (pos, neg) score vectors incl. heavy ties / clamp plateaus / extreme imbalance."""
    cases = {}
    cases["kat1"] = (np.float32([.9, .8, .8, .35, .2]), np.float32([.8, .4, .35, .1, .1, .05]))
    rng = np.random.default_rng(1234)
    s = (rng.integers(0, 64, 4096) / 63).astype(np.float32)
    y = rng.random(4096) < 0.1 + 0.5 * s
    cases["kat2"] = (s[y], s[~y])
    s = rng.standard_normal(100000).astype(np.float32)
    y = rng.random(100000) < 0.03
    s[y] += 1.5
    cases["kat3"] = (s[y], s[~y])
    cases["kat4"] = (np.float32([3, 2]), np.float32([1, 0]))
    cases["kat5"] = (np.float32([1, 1]), np.float32([2, .5, .2]))
    # clamp plateau: 30% of all scores sit on one value (as after the 400 clamp), scores in [-1, 0]
    s = -rng.random(50000).astype(np.float32)
    y = rng.random(50000) < 0.02
    s[y] *= 0.5
    plateau = rng.random(50000) < 0.3
    s[plateau] = -1.0
    cases["plateau"] = (s[y], s[~y])
    # all scores identical
    cases["allties"] = (np.full(7, 0.25, np.float32), np.full(11, 0.25, np.float32))
    # single positive / single negative
    s = rng.standard_normal(1000).astype(np.float32)
    cases["onepos"] = (s[:1], s[1:])
    cases["oneneg"] = (s[1:], s[:1])
    # mixed-sign, wide dynamic range, with +-0
    s = (rng.standard_normal(20000) * np.exp(rng.standard_normal(20000) * 4)).astype(np.float32)
    s[:50] = 0.0
    s[50:100] = -0.0
    y = rng.random(20000) < 0.2
    cases["widerange"] = (s[y], s[~y])
    # recall ties around 0.95: 20 positives -> recall steps of 0.05
    s = rng.random(400).astype(np.float32)
    y = np.zeros(400, bool)
    y[:20] = True
    cases["recallsteps"] = (s[y], s[~y])
    return cases


def gen_metrics():
    out = {}
    with reference("anomaly"):
        import anom_utils
        import eval_ood_traditional as E
        for name, (pos, neg) in metric_cases().items():
            auroc, aupr, fpr = anom_utils.get_measures(pos, neg)
            out[f"{name}_pos"] = pos
            out[f"{name}_neg"] = neg
            out[f"{name}_res"] = np.float64([auroc, aupr, fpr])
            labels = np.r_[np.ones(len(pos), np.int32), np.zeros(len(neg), np.int32)]
            ex = np.r_[pos, neg]
            out[f"{name}_fpr90"] = np.float64(anom_utils.fpr_and_fdr_at_recall(labels, ex, 0.90))
        # image-level wrappers
        rng = np.random.default_rng(7)
        conf = rng.random((48, 64)).astype(np.float32)
        seg = rng.integers(0, 13, (48, 64)).astype(np.int64)
        seg[10:20, 10:30] = 13
        conf[10:20, 10:30] *= 0.6
        seg[30:34, 40:50] = -1
        cfg = CfgNode(OOD=CfgNode(out_labels=(13,)))
        out["img_conf"] = conf
        out["img_seg"] = seg
        out["img_res_script"] = np.float64(E.eval_ood_measure(conf, seg, cfg))
        out["img_res_anom_utils"] = np.float64(anom_utils.eval_ood_measure(conf, seg, out_label=13))
        cfg2 = CfgNode(OOD=CfgNode(out_labels=(13, 5)))
        out["img_res_script_two_labels"] = np.float64(E.eval_ood_measure(conf, seg, cfg2))
        mask = rng.random((48, 64)) < 0.7
        out["img_mask"] = mask
        # reference quirk: `mask` filters seg_label only (eval_ood_traditional.py:130-131); conf must arrive pre-masked
        out["img_res_script_masked"] = np.float64(E.eval_ood_measure(conf[mask], seg, cfg, mask=mask))
        with contextlib.redirect_stdout(io.StringIO()):
            none_res = E.eval_ood_measure(conf, np.zeros_like(seg), cfg)
        assert none_res is None
        # helper functions
        x = rng.standard_normal((16, 16)).astype(np.float32) * 100
        out["norm_in"] = x
        out["norm_out"] = E.Normalizatoin(x)
        e = rng.random((16, 16)).astype(np.float32)
        out["coef_in"] = e
        out["coef_out"] = E.Coefficient_map(e, 0.2)
    save("metrics_kat.npz", **out)


# --------------------------------------------------------------------------- #
def gen_head_anomaly():
    out = {}
    with reference("anomaly"), contextlib.redirect_stdout(io.StringIO()):
        from models import models as M
        torch.manual_seed(11)
        for tag, scale in (("a", 1.0), ("b", 6.0)):
            dec = M.PPMDeepsup_embedding(num_class=13, fc_dim=32, use_softmax=True).eval()
            dec.conv_last[4].weight.data.mul_(scale)
            captured = {}
            dec.conv_last.register_forward_hook(lambda m, i, o: captured.__setitem__("x", o.detach().clone()))
            conv_out = [torch.randn(1, 16, 9, 12), torch.randn(1, 32, 9, 12)]
            with torch.no_grad():
                z_up, f_up = dec(conv_out, segSize=(40, 56))
                z_only = dec(conv_out, segSize=(40, 56), output_ft=False)
            out[f"{tag}_x_low"] = captured["x"].numpy()
            out[f"{tag}_z_up"] = z_up.numpy()
            out[f"{tag}_f_up"] = f_up.numpy()
            assert torch.equal(z_only, z_up)
            # train-mode return: ((logits_lowres, deepsup), ft)
            dec.use_softmax = False
            with torch.no_grad():
                (z_low, deepsup), ft = dec(conv_out)
            out[f"{tag}_z_low"] = z_low.numpy()
            out[f"{tag}_centers"] = dec.centers.numpy()
    save("head_anomaly.npz", **out)


def gen_head_deeplab():
    out = {}
    with reference("DeepLabV3Plus-Pytorch"), contextlib.redirect_stdout(io.StringIO()):
        from network import utils as NU
        torch.manual_seed(12)

        class TinyBackbone(nn.Module):
            def forward(self, x):
                return x

        for k in (16, 17, 19):
            cls = nn.Conv2d(3, k, 3, padding=1, stride=2)
            cls.weight.data.mul_(8.0)
            model = NU._SimpleSegmentationModel_embedding(TinyBackbone(), cls).eval()
            img = torch.randn(2, 3, 24, 40)
            with torch.no_grad():
                logits, centers, feats = model(img)
            out[f"k{k}_features_nhwc"] = feats.numpy()
            out[f"k{k}_logits"] = logits.numpy()
            out[f"k{k}_centers"] = centers.numpy()
    save("head_deeplab.npz", **out)


# --------------------------------------------------------------------------- #
def gen_evaluate_anomaly():
    """Run the reference's evaluate() end to end (config 1 shape-reduced: ResNet18-dilated
    PSPNet, 2 images, 5 scales) and capture its intermediate and final values."""
    out = {}
    with reference("anomaly"):
        import eval_ood_traditional as E
        from models import models as M
        from models import resnet
        torch.manual_seed(21)
        with contextlib.redirect_stdout(io.StringIO()):
            enc = M.ResnetDilated(resnet.resnet18(pretrained=False), dilate_scale=8)
            dec = M.PPMDeepsup_embedding(num_class=13, fc_dim=512, use_softmax=True)
        dec.conv_last[4].weight.data.mul_(8.0)
        dec.conv_last[4].bias.data.zero_()
        module = M.SegmentationModule(enc, dec, nn.NLLLoss(ignore_index=-1))

        lows, calls, preds = [], [], []
        dec.conv_last.register_forward_hook(lambda m, i, o: lows.append(o.detach().clone().numpy()))
        ref_measure = E.eval_ood_measure

        def spy_measure(conf, seg_label, cfg, mask=None):
            res = ref_measure(conf, seg_label, cfg, mask=mask)
            calls.append((np.array(conf), np.array(seg_label), res))
            return res

        ref_acc, ref_iu = E.accuracy, E.intersectionAndUnion
        accs, ius = [], []

        def spy_acc(pred, label):
            preds.append(np.array(pred))
            r = ref_acc(pred, label)
            accs.append(r)
            return r

        def spy_iu(pred, label, n):
            r = ref_iu(pred, label, n)
            ius.append(r)
            return r

        E.eval_ood_measure, E.accuracy, E.intersectionAndUnion = spy_measure, spy_acc, spy_iu
        H, W = 96, 160
        sizes = [(40, 72), (56, 88), (64, 104), (72, 120), (80, 136)]
        cfg = CfgNode(DATASET=CfgNode(num_class=13, imgSizes=(1, 2, 3, 4, 5)),
                      OOD=CfgNode(exclude_back=False, ood="dissum", out_labels=(13,)),
                      VAL=CfgNode(visualize=False), DIR="/tmp")
        loader = []
        g = torch.Generator().manual_seed(22)
        for i in range(2):
            seg = torch.randint(0, 13, (1, H, W), generator=g)
            seg[0, 20:50, 30:90] = 13
            seg[0, 0:4, :] = -1
            loader.append([{"img_ori": np.zeros((H, W, 3), np.uint8),
                            "img_data": [torch.randn(1, 3, h, w, generator=g) for h, w in sizes],
                            "seg_label": seg, "info": f"img{i}.jpg", "name": f"img{i}"}])
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
            E.evaluate(module, loader, cfg, 0)
        summary = [ln for ln in buf.getvalue().splitlines() if "mean auroc" in ln or "Mean IoU" in ln]
        print("\n".join(summary))
        assert len(lows) == 10 and len(calls) == 2 and len(preds) == 2
        for i in range(2):
            for s in range(5):
                out[f"img{i}_low{s}"] = lows[i * 5 + s]
            conf, seg, res = calls[i]
            out[f"img{i}_conf"] = conf
            out[f"img{i}_seg"] = seg
            out[f"img{i}_res"] = np.float64(res)
            out[f"img{i}_pred"] = preds[i]
            out[f"img{i}_acc"] = np.float64([accs[i][0], accs[i][1]])
            out[f"img{i}_inter"] = ius[i][0]
            out[f"img{i}_union"] = ius[i][1]
        out["summary"] = np.array(summary)
    save("evaluate_anomaly.npz", **out)


def gen_evaluate_anomaly_modes():
    """The reference's evaluate() in its other score modes -- `--ood msp`, `--ood maxlogit` and `OOD.exclude_back`
    (anomaly/eval_ood_traditional.py:212-214,276-278,288-290,302-305) -- one image per mode, same reduced model as
    gen_evaluate_anomaly(): conf map handed to eval_ood_measure, its result, pred and the stride-8 embeddings."""
    out = {}
    modes = [("msp", False), ("maxlogit", False), ("dissum", True), ("msp", True), ("maxlogit", True)]
    with reference("anomaly"):
        import eval_ood_traditional as E
        from models import models as M
        from models import resnet
        torch.manual_seed(31)
        with contextlib.redirect_stdout(io.StringIO()):
            enc = M.ResnetDilated(resnet.resnet18(pretrained=False), dilate_scale=8)
            dec = M.PPMDeepsup_embedding(num_class=13, fc_dim=512, use_softmax=True)
        dec.conv_last[4].weight.data.mul_(8.0)
        dec.conv_last[4].bias.data.zero_()
        module = M.SegmentationModule(enc, dec, nn.NLLLoss(ignore_index=-1))
        lows, calls, preds = [], [], []
        dec.conv_last.register_forward_hook(lambda m, i, o: lows.append(o.detach().clone().numpy()))
        ref_measure, ref_acc = E.eval_ood_measure, E.accuracy

        def spy_measure(conf, seg_label, cfg, mask=None):
            res = ref_measure(conf, seg_label, cfg, mask=mask)
            calls.append((np.array(conf), np.array(seg_label), res))
            return res

        def spy_acc(pred, label):
            preds.append(np.array(pred))
            return ref_acc(pred, label)

        E.eval_ood_measure, E.accuracy = spy_measure, spy_acc
        try:
            H, W = 96, 160
            sizes = [(40, 72), (56, 88), (64, 104), (72, 120), (80, 136)]
            g = torch.Generator().manual_seed(32)
            for mi, (mode, excl) in enumerate(modes):
                cfg = CfgNode(DATASET=CfgNode(num_class=13, imgSizes=(1, 2, 3, 4, 5)),
                              OOD=CfgNode(exclude_back=excl, ood=mode, out_labels=(13,)),
                              VAL=CfgNode(visualize=False), DIR="/tmp")
                seg = torch.randint(0, 13, (1, H, W), generator=g)
                seg[0, 30:60, 40:110] = 13
                seg[0, 0:3, :] = -1
                loader = [[{"img_ori": np.zeros((H, W, 3), np.uint8),
                            "img_data": [torch.randn(1, 3, h, w, generator=g) for h, w in sizes],
                            "seg_label": seg, "info": f"m{mi}.jpg", "name": f"m{mi}"}]]
                n0 = len(calls)
                with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                    E.evaluate(module, loader, cfg, 0)
                assert len(calls) == n0 + 1 and len(lows) == 5 * (mi + 1) and len(preds) == mi + 1
                tag = f"{mode}_{'noback' if excl else 'all'}"
                for sidx in range(5):
                    out[f"{tag}_low{sidx}"] = lows[5 * mi + sidx]
                conf, sg, res = calls[-1]
                out[f"{tag}_conf"], out[f"{tag}_seg"], out[f"{tag}_res"] = conf, sg, np.float64(res)
                out[f"{tag}_pred"] = preds[-1]
        finally:
            E.eval_ood_measure, E.accuracy = ref_measure, ref_acc
    save("evaluate_anomaly_modes.npz", **out)


def gen_config0_full_shape():
    """BASELINE.json configs[0] at its real shape: the reference's evaluate() (`--ood dissum`) with a random-init
    PSPNet-ResNet50dilated embedding model on 4 synthetic 720x1280 StreetHazards-shape images (5 scales), on the CPU.
    Kept small: the stride-8 embeddings, pred and a subsampled conf map of image 0 only; results of all 4 images."""
    out = {}
    with reference("anomaly"):
        import eval_ood_traditional as E
        from models import models as M
        from models import resnet
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            enc = M.ResnetDilated(resnet.resnet50(pretrained=False), dilate_scale=8)
            dec = M.PPMDeepsup_embedding(num_class=13, fc_dim=2048, use_softmax=True)
        dec.conv_last[4].weight.data.mul_(0.3)       # random init gives |x|^2 ~ 196 (sum d ~ 2800 >= the 400 clamp everywhere -> 0/0);
        # 0.3 puts the median sum d near 350 with a few % of the pixels on the clamp plateau
        dec.conv_last[4].bias.data.zero_()
        module = M.SegmentationModule(enc, dec, nn.NLLLoss(ignore_index=-1))
        lows, calls, preds, accs, ius = [], [], [], [], []
        dec.conv_last.register_forward_hook(lambda m, i, o: lows.append(o.detach().clone().numpy()))
        ref_measure, ref_acc, ref_iu = E.eval_ood_measure, E.accuracy, E.intersectionAndUnion

        def spy_measure(conf, seg_label, cfg, mask=None):
            res = ref_measure(conf, seg_label, cfg, mask=mask)
            calls.append((np.array(conf), np.array(seg_label), res))
            return res

        def spy_acc(pred, label):
            preds.append(np.array(pred))
            r = ref_acc(pred, label)
            accs.append(r)
            return r

        def spy_iu(pred, label, n):
            r = ref_iu(pred, label, n)
            ius.append(r)
            return r

        E.eval_ood_measure, E.accuracy, E.intersectionAndUnion = spy_measure, spy_acc, spy_iu
        try:
            H, W = 720, 1280
            sizes = [(304, 536), (376, 672), (456, 800), (528, 936), (568, 1000)]      # anomaly/dataset.py:281-289 on 720x1280
            cfg = CfgNode(DATASET=CfgNode(num_class=13, imgSizes=(1, 2, 3, 4, 5)),
                          OOD=CfgNode(exclude_back=False, ood="dissum", out_labels=(13,)),
                          VAL=CfgNode(visualize=False), DIR="/tmp")
            g = torch.Generator().manual_seed(1)
            loader = []
            for i in range(4):
                seg = torch.randint(0, 13, (1, H, W), generator=g)
                seg[0, 100 + 50 * i:200 + 50 * i, 300:500] = 13
                loader.append([{"img_ori": np.zeros((H, W, 3), np.uint8),
                                "img_data": [torch.randn(1, 3, h, w, generator=g) for h, w in sizes],
                                "seg_label": seg, "info": f"c0_{i}.jpg", "name": f"c0_{i}"}])
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
                E.evaluate(module, loader, cfg, 0)
        finally:
            E.eval_ood_measure, E.accuracy, E.intersectionAndUnion = ref_measure, ref_acc, ref_iu
        assert len(lows) == 20 and len(calls) == 4 and len(preds) == 4
        for s_ in range(5):
            out[f"img0_low{s_}"] = lows[s_]
        out["img0_pred"] = preds[0].astype(np.uint8)
        out["img0_conf_sub8"] = calls[0][0][::8, ::8].copy()
        out["img0_seg"] = calls[0][1].astype(np.int16)
        for i in range(4):
            conf, seg, res = calls[i]
            out[f"img{i}_res"] = np.float64(res)
            out[f"img{i}_conf_sum"] = np.float64(conf.astype(np.float64).sum())
            out[f"img{i}_conf_minmax_raw"] = np.float32([conf.min(), conf.max()])
            out[f"img{i}_pred_hist"] = np.bincount(preds[i].reshape(-1).astype(np.int64), minlength=13)
            out[f"img{i}_acc"] = np.float64([accs[i][0], accs[i][1]])
            out[f"img{i}_inter"], out[f"img{i}_union"] = ius[i][0], ius[i][1]
        out["summary"] = np.array([ln for ln in buf.getvalue().splitlines() if "mean auroc" in ln or "Mean IoU" in ln])
    save("config0_full_shape.npz", **out)


def gen_validate_deeplab():
    """Run the reference's test_embedding.validate() on a synthetic loader (batch 1) with a tiny
    backbone and capture preds / confusion matrix."""
    out = {}
    with reference("DeepLabV3Plus-Pytorch") as root:
        import tempfile
        tmp = tempfile.mkdtemp()
        os.chdir(tmp)
        torch.manual_seed(31)
        rng = np.random.default_rng(32)
        with contextlib.redirect_stdout(io.StringIO()):
            import test_embedding as T
            from network import utils as NU
            from metrics import StreamSegMetrics

        class TinyBackbone(nn.Module):
            def forward(self, x):
                return x

        cls = nn.Conv2d(3, 16, 3, padding=1, stride=2)
        cls.weight.data.mul_(0.7)
        model = NU._SimpleSegmentationModel_embedding(TinyBackbone(), cls).eval()
        images = [torch.randn(1, 3, 32, 48) for _ in range(3)]
        # place the novel prototype near where real features live so that the override fires
        with torch.no_grad():
            _, _, f0 = model(images[0])
        f0 = f0.reshape(-1, 16).numpy()
        base = f0[rng.integers(0, len(f0), 5)].astype(np.float64)
        protos = (base + 0.05 * rng.standard_normal(base.shape)).tolist()
        with open("prototype_car_5_shot.json", "w") as fh:
            json.dump(protos, fh)
        loader = []
        for img in images:
            lab = torch.from_numpy(rng.integers(0, 19, (1, 32, 48))).long()
            lab[0, :2, :] = 255
            lab[0, 2, :] = 254
            loader.append((img, lab, lab.clone()))
        labels_in = [lab.clone() for _, lab, _ in loader]   # validate() remaps `labels` in place
        metrics = StreamSegMetrics(16)
        captured = []
        ref_update = metrics.update

        def spy_update(t, p):
            captured.append((np.array(t), np.array(p)))
            return ref_update(t, p)

        metrics.update = spy_update
        opts = SimpleNamespace(save_val_results=False, num_classes=16)
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            res = T.validate(opts, model, loader, torch.device("cpu"), metrics)
        score = res[0] if isinstance(res, tuple) else res
        out["prototypes"] = np.float64(protos)
        for i, (img, lab, _) in enumerate(loader):
            with torch.no_grad():
                logits, _, feats = model(img)
            out[f"img{i}_features_nhwc"] = feats.numpy()
            out[f"img{i}_logits"] = logits.numpy()
            out[f"img{i}_labels_in"] = labels_in[i].numpy()
            out[f"img{i}_targets"] = captured[i][0]
            out[f"img{i}_preds"] = captured[i][1]
            print("novel pixels:", int((captured[i][1] == 16).sum()))
        out["confusion"] = metrics.confusion_matrix
        for key in ("Overall Acc", "Mean Acc", "FreqW Acc", "Mean IoU"):
            out["res_" + key.replace(" ", "_")] = np.float64(score[key])
        out["res_class_iou"] = np.float64([score["Class IoU"][k] for k in range(19)])
        # Normalization helper + score block pieces replayed by validate (test_embedding.py:150-153)
        x = rng.standard_normal((8, 8)).astype(np.float32)
        out["norm_in"], out["norm_out"] = x, T.Normalization(x)
        os.chdir(root)
    save("validate_deeplab.npz", **out)


def gen_plm():
    """PLM head (self-distillation model) forward + the merge lines of its validate()."""
    out = {}
    with reference("DeepLabV3Plus-Pytorch"), contextlib.redirect_stdout(io.StringIO()):
        from network import utils as NU
        torch.manual_seed(41)

        class Net(NU._SimpleSegmentationModel_embedding_self_distillation):
            """Same forward()/forward_single(); only the heavy DeepLab heads are swapped for 1x1 convs."""

            def __init__(self):
                nn.Module.__init__(self)
                self.backbone = nn.Identity()
                self.classifier_list = ["classifier", "classifier_1"]
                self.cls_novel = 1
                self.classifier = nn.Conv2d(3, 16, 1)
                self.classifier_1 = nn.Conv2d(3, 17, 1)
                self.centers = torch.zeros(17, 17)

        net = Net().eval()
        net.classifier.weight.data.mul_(5)
        net.classifier_1.weight.data.mul_(5)
        img = torch.randn(2, 3, 16, 24)
        with torch.no_grad():
            logits, centers, feats = net(img)
        for i in range(2):
            out[f"head{i}_features_nhwc"] = feats[i].numpy()
            out[f"head{i}_logits"] = logits[i].numpy()
            out[f"head{i}_centers"] = centers[i].numpy()
        # test_self_distillation.py:292-297 executed literally on these outputs
        outputs = logits
        opts = SimpleNamespace(novel_cls=1)
        preds_base = outputs[0].detach().max(dim=1)[1]
        for i in range(opts.novel_cls):
            labels_base = outputs[i + 1].detach().max(dim=1)[1]
            preds_base[labels_base == (16 + i)] = 16 + i
        out["merged_preds"] = preds_base.numpy()
    save("plm.npz", **out)


# --------------------------------------------------------------------------- #
def gen_loss():
    out = {}
    torch.manual_seed(51)
    n, k, h, w = 2, 16, 12, 20
    emb = torch.randn(n, k, h, w) * 1.5
    target = torch.randint(0, k, (n, h, w))
    target[0, :3, :] = 255
    target[1, 5, 2:9] = 255
    with reference("DeepLabV3Plus-Pytorch"), contextlib.redirect_stdout(io.StringIO()):
        from network import utils as NU
        import utils.loss as L

        def logits_of(e):
            # same distance math as the reference head, differentiable
            f = e.permute(0, 2, 3, 1).contiguous().view(n, h * w, k)
            f = f.unsqueeze(2).expand(n, h * w, k, k)
            c = torch.zeros(k, k)
            for i in range(k):
                c[i][i] = 3
            return (-torch.sum((f - c) ** 2, 3)).permute(0, 2, 1).contiguous().view(n, k, h, w)

        # shipped state: CE / n
        e1 = emb.clone().requires_grad_(True)
        z = logits_of(e1)
        crit = L.CrossEntropyLoss(alpha=0.01, beta=0.01 / 80, gamma=0)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            loss = crit(z, target, None)
        loss.backward()
        out["emb"], out["target"] = emb.numpy(), target.numpy()
        out["shipped_loss"] = loss.detach().numpy()
        out["shipped_grad_emb"] = e1.grad.numpy()
        out["logits"] = z.detach().numpy()

        # intended full form (utils/loss.py:43-79): the reference's own source with ONLY the early
        # `return CE_loss / n` line removed, executed here; features passed as [n, HW, C] so that the
        # flat-pixel index_select of :53,65 addresses pixels (the evident intent).
        src = textwrap.dedent(inspect.getsource(L.CrossEntropyLoss.forward))
        lines = src.splitlines()
        idx = [i for i, ln in enumerate(lines) if ln.strip() == "return CE_loss / n"]
        assert len(idx) == 1
        del lines[idx[0]]
        ns = dict(nn=nn, torch=torch, np=np, Variable=torch.autograd.Variable)
        exec("\n".join(lines), ns)
        full_forward = ns["forward"]
        for tag, (a, b, g) in {"abg": (0.01, 0.01 / 80, 0.0), "vl": (0.01, 0.0, 0.0), "center": (0.01, 0.0, 0.05)}.items():
            crit = L.CrossEntropyLoss(alpha=a, beta=b, gamma=g)
            e2 = emb.clone().requires_grad_(True)
            z2 = logits_of(e2)
            feats = e2.permute(0, 2, 3, 1).contiguous().view(n, h * w, k)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                loss2 = full_forward(crit, z2, target, feats)
            loss2.backward()
            out[f"full_{tag}_coef"] = np.float64([a, b, g])
            out[f"full_{tag}_loss"] = loss2.detach().numpy().reshape(())
            out[f"full_{tag}_grad_emb"] = e2.grad.numpy()

    # anomaly train branch (models.py:36-84): CE(ignore -1) + alpha * VL, alpha = 0.01
    with reference("anomaly"), contextlib.redirect_stdout(io.StringIO()):
        from models import models as M
        k13 = 13
        emb13 = (torch.randn(2, k13, 10, 14) * 1.5)
        tgt13 = torch.randint(0, k13, (2, 10, 14))
        tgt13[0, :2, :] = -1
        e3 = emb13.clone().requires_grad_(True)

        class Enc(nn.Module):
            def forward(self, x, return_feature_maps=True):
                return [x]

        class Dec(nn.Module):
            """Stands in for conv layers only; the distance block is the reference's math."""

            def forward(self, conv_out, segSize=None):
                x = conv_out[-1]
                f = x.permute(0, 2, 3, 1).contiguous()
                f = f.view(f.shape[0], -1, f.shape[3]).unsqueeze(2).expand(-1, -1, k13, -1)
                c = torch.zeros(k13, k13)
                for i in range(k13):
                    c[i][i] = 3
                d = -torch.sum((f - c) ** 2, 3)
                return d.permute(0, 2, 1).contiguous().view(x.shape[0], k13, x.shape[2], x.shape[3])

        mod = M.SegmentationModule(Enc(), Dec(), nn.CrossEntropyLoss(ignore_index=-1))
        loss3, acc3 = mod({"img_data": e3, "seg_label": tgt13})
        loss3.backward()
        out["anom_emb"], out["anom_target"] = emb13.numpy(), tgt13.numpy()
        out["anom_loss"] = loss3.detach().numpy().reshape(())
        out["anom_acc"] = acc3.detach().numpy()
        out["anom_grad_emb"] = e3.grad.numpy()
    save("loss.npz", **out)


def gen_segmetrics():
    out = {}
    rng = np.random.default_rng(61)
    with reference("DeepLabV3Plus-Pytorch"), contextlib.redirect_stdout(io.StringIO()):
        from metrics import StreamSegMetrics
        m = StreamSegMetrics(16)
        m.reset()
        gts, prs = [], []
        for _ in range(3):
            gt = rng.integers(0, 19, (2, 20, 30)).astype(np.int64)
            gt[:, :2] = 255
            pr = np.where(rng.random(gt.shape) < 0.7, np.minimum(gt, 16), rng.integers(0, 17, gt.shape)).astype(np.int64)
            m.update(gt, pr)
            gts.append(gt)
            prs.append(pr)
        res = m.get_results()
        out["dl_gt"], out["dl_pred"] = np.stack(gts), np.stack(prs)
        out["dl_confusion"] = m.confusion_matrix
        for key in ("Overall Acc", "Mean Acc", "FreqW Acc", "Mean IoU"):
            out["dl_" + key.replace(" ", "_")] = np.float64(res[key])
        out["dl_class_iou"] = np.float64([res["Class IoU"][k] for k in range(19)])
        out["dl_to_str"] = np.array(m.to_str(res))
    with reference("anomaly"), contextlib.redirect_stdout(io.StringIO()):
        from utils import accuracy, intersectionAndUnion
        gt = rng.integers(-1, 14, (40, 50)).astype(np.int64)
        pr = np.where(rng.random(gt.shape) < 0.6, np.clip(gt, 0, 12), rng.integers(0, 13, gt.shape)).astype(np.int64)
        acc, pix = accuracy(pr, gt)
        inter, union = intersectionAndUnion(pr, gt, 13)
        out["an_gt"], out["an_pred"] = gt, pr
        out["an_acc"] = np.float64([acc, pix])
        out["an_inter"], out["an_union"] = inter, union
    save("segmetrics.npz", **out)


def gen_roc_baseline():
    """Softmax-baseline evaluator of DeepLabV3Plus-Pytorch/test.py:241-244, which calls scikit-learn directly
    (``import sklearn.metrics as Metrics``): roc_auc_score, roc_curve (drop_intermediate default),
    average_precision_score, fpr95 = fpr[tpr >= 0.95][0].  Cases stress the drop_intermediate rule: runs of
    consecutive groups with identical (pos, neg) counts around the 95 % recall point, ties, plateaus."""
    import sklearn.metrics as Metrics
    rng = np.random.default_rng(77)
    cases = {}
    # continuous scores (every group has one element)
    s = rng.random(20000).astype(np.float32)
    y = (rng.random(20000) < 0.05 + 0.3 * s).astype(np.int64)
    cases["continuous"] = (y, s)
    # coarse quantisation: large groups, many with equal counts
    s = (rng.integers(0, 40, 30000) / 40).astype(np.float32)
    y = (rng.random(30000) < 0.2).astype(np.int64)
    cases["quantised"] = (y, s)
    # pairs (1 pos, 1 neg) per score value over a long stretch covering the recall point: collinear diagonal
    k = 2000
    s = np.repeat(np.arange(k, dtype=np.float32) / k, 2)
    y = np.tile(np.array([1, 0]), k)
    cases["diagonal_pairs"] = (y, s)
    # same, but the stretch ends right after the recall point
    s2 = s.copy(); y2 = y.copy()
    y2[:150] = 0
    cases["diagonal_then_negatives"] = (y2, s2)
    # plateau at the top score holding > 95 % of the positives (first point already reaches the level)
    s = rng.random(5000).astype(np.float32) * 0.5
    y = np.zeros(5000, np.int64)
    s[:300] = 0.9; y[:290] = 1; y[4000:4010] = 1
    cases["plateau_first_point"] = (y, s)
    # tiny
    cases["tiny"] = (np.array([1, 0, 1, 1, 0, 0, 1, 0]), np.float32([.9, .8, .8, .7, .7, .3, .2, .1]))
    out = {}
    for name, (y, s) in cases.items():
        auc = Metrics.roc_auc_score(y, s)
        fpr, tpr, _ = Metrics.roc_curve(y, s)
        aupr = Metrics.average_precision_score(y, s)
        out[f"{name}_y"] = y.astype(np.uint8)
        out[f"{name}_s"] = s.astype(np.float32)
        out[f"{name}_res"] = np.float64([auc, aupr, fpr[tpr >= 0.95][0]])
        out[f"{name}_fpr90"] = np.float64(fpr[tpr >= 0.90][0])
    save("roc_baseline.npz", **out)


def main():
    import sklearn
    import scipy
    if len(sys.argv) > 1:                      # regenerate only the named fixtures: make_golden.py gen_evaluate_anomaly_modes ...
        for name in sys.argv[1:]:
            globals()[name]()
        return
    gen_evaluate_anomaly_modes()
    gen_config0_full_shape()
    gen_metrics()
    gen_head_anomaly()
    gen_head_deeplab()
    gen_evaluate_anomaly()
    gen_validate_deeplab()
    gen_plm()
    gen_loss()
    gen_segmetrics()
    gen_roc_baseline()
    meta = {"python": sys.version.split()[0], "torch": torch.__version__, "numpy": np.__version__,
            "sklearn": sklearn.__version__, "scipy": scipy.__version__,
            "reference": "/root/reference (Jun-CEN/Open-World-Semantic-Segmentation, unmodified)"}
    with open(os.path.join(HERE, "meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print(meta)


if __name__ == "__main__":
    main()
