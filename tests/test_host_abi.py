"""CPU tier: the C-ABI library loads without a GPU and exports every symbol include/dml_b200.h declares;
the ctypes binding covers exactly those symbols; host-side helpers behave; the product refuses to run
without CUDA (no CPU fallback).  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dml_b200.h")).read()
    return sorted(set(re.findall(r"DML_API\s+[\w\s\*]+?\b(dml_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import dml_b200
    lib = dml_b200.load_library()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libdml_b200.so does not export {n}"
    assert lib.dml_abi_version() == 4
    assert lib.dml_max_dim() == 32
    assert lib.dml_error_string(-2).decode().startswith("embedding dim")


def test_binding_covers_header_exactly():
    from dml_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_struct_layouts_match_the_header():
    from dml_b200 import _lib
    # dml_head_params: field order/types as declared (sizes follow the platform ABI)
    text = open(os.path.join(ROOT, "include", "dml_b200.h")).read()
    start = text.index("typedef struct dml_head_params {") + len("typedef struct dml_head_params {")
    body = text[start:text.index("} dml_head_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(",")
        first = names[0].split()
        fields.append(first[-1].lstrip("*"))
        fields.extend(n.strip().lstrip("*") for n in names[1:])
    assert [f[0] for f in _lib.HeadParams._fields_] == fields
    assert ctypes.sizeof(_lib.OodResult) == 56


def test_multiscale_struct_layout_matches_the_header():
    from dml_b200 import _lib
    text = open(os.path.join(ROOT, "include", "dml_b200.h")).read()
    start = text.index("typedef struct dml_multiscale_params {") + len("typedef struct dml_multiscale_params {")
    body = re.sub(r"/\*.*?\*/", "", text[start:text.index("} dml_multiscale_params;")], flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(",")
        fields.append(names[0].split()[-1].lstrip("*"))
        fields.extend(n.strip().lstrip("*") for n in names[1:])
    fields = [re.sub(r"\[.*\]", "", f) for f in fields]
    assert [f[0] for f in _lib.MultiscaleParams._fields_] == fields
    assert _lib.MAX_SCALES == int(re.search(r"#define DML_MAX_SCALES (\d+)", text).group(1))
    # 5 x i32 + pad, 8 pointers, 16 x i32, 3 x 4-byte scalars + pad, 7 pointers, 2 x i32, 3 pointers, 2 x i32
    assert ctypes.sizeof(_lib.MultiscaleParams) == 24 + 64 + 64 + 16 + 56 + 8 + 24 + 8


def test_no_cpu_fallback():
    import dml_b200
    with pytest.raises(dml_b200.DmlError):
        dml_b200.dml_head(torch.zeros(1, 13, 4, 4))
    with pytest.raises(dml_b200.DmlError):
        dml_b200.dml_multiscale_head([torch.zeros(1, 13, 4, 4)], (8, 8))
    from dml_b200 import ood
    with pytest.raises(dml_b200.DmlError):
        ood.eval_segments(torch.zeros(8), 1, 8, gt=torch.zeros(8, dtype=torch.uint8))
    with pytest.raises(dml_b200.DmlError):
        dml_b200.dml_loss(torch.zeros(1, 13, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64))


def test_missing_library_fails_loudly(tmp_path):
    from dml_b200 import _lib
    with pytest.raises(_lib.DmlError):
        _lib.load_library(str(tmp_path / "nope.so"))


def test_host_helpers():
    from dml_b200 import head, ood
    assert ood.label_mask((13,)) == 1 << 13 and ood.label_mask((13, 5)) == (1 << 13) | (1 << 5)
    with pytest.raises(ValueError):
        ood.label_mask((64,))
    assert head.scaled_identity_magnitude(torch.eye(13) * 3) == 3.0
    assert head.scaled_identity_magnitude(torch.ones(3, 3)) is None
    assert head.scaled_identity_magnitude(torch.eye(3)[:2]) is None
    # combining range partials: exact sums, two integer FPR candidates, float64 choice with "later group wins" ties
    def part(num, ap, a, b, groups):
        p = np.zeros(10, np.int64)
        p.view(np.uint64)[0] = num
        p.view(np.float64)[1] = ap
        p[2:5] = a
        p[5:8] = b
        p[8] = groups
        return p
    none_b = (ood.NO_B, -1, 0)
    p0 = part(10, 1.5, (7, 18, 3), none_b, 4)          # a: idx 7, tps 18 (recall 0.90), fps 3
    p1 = part(6, 0.5, (-1, 0, 0), (20, 9, 5), 2)       # b: tps 20 (recall 1.00), idx 9, fps 5
    a, p, f, g = ood.combine_partials([p0, p1], total_pos=20, total_n=28, recall_level=0.95)
    assert a == 16 / (2.0 * 20 * 8) and p == 2.0 / 20 and g == 6
    # |0.90-0.95| = 0.04999999999999993 < |1.00-0.95| = 0.050000000000000044 in float64 -> candidate a
    assert f == 3 / 8
    assert ood.combine_partials([p0, p1], 20, 28, recall_level=0.96)[2] == 5 / 8
    assert np.isnan(ood.combine_partials([p0], 0, 12)[0])


def test_drop_in_surface_names():
    import dml_b200
    from dml_b200.anomaly import anom_utils, eval_ood, models, utils
    from dml_b200 import deeplab
    for fn in ("get_measures", "fpr_and_fdr_at_recall", "get_and_print_results", "eval_ood_measure", "print_measures",
               "print_measures_with_std"):
        assert callable(getattr(anom_utils, fn))
    assert anom_utils.recall_level_default == 0.95
    for fn in ("accuracy", "intersectionAndUnion"):
        assert callable(getattr(utils, fn))
    assert callable(eval_ood.eval_ood_measure) and callable(eval_ood.score_map)
    m = models.PPMDeepsup_embedding(num_class=13, fc_dim=64, use_softmax=True)
    assert torch.equal(m.centers, torch.eye(13) * 3)
    for name in ("_SimpleSegmentationModel_embedding", "_SimpleSegmentationModel_embedding_self_distillation",
                 "CrossEntropyLoss", "CrossEntropyLoss_dis", "FocalLoss", "StreamSegMetrics"):
        assert hasattr(deeplab, name)
    s = deeplab.StreamSegMetrics(16)
    assert s.n_classes == 19 and s.confusion_matrix.shape == (16, 16)
    s.reset()
    assert s.confusion_matrix.shape == (19, 19)
    crit = deeplab.CrossEntropyLoss(alpha=0.01, beta=0.01 / 80, gamma=0)
    assert crit.ignore_index == 255 and crit.shipped_early_return
