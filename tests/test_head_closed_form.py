"""CPU tier: the closed forms the lean scoring path of head_kernel<D, IDENT, VEC, false> evaluates for prototypes
m*I (csrc/dml_head.cuh) -- restated in NumPy float32 with the kernel's operation order -- against the oracle's direct
form (DeepLabV3Plus-Pytorch/network/utils.py:98-111 / anomaly/models/models.py:645-651):

    sum_k d_k      = (D-1) S + sum_k (x_k - m)^2                    S = sum_k x_k^2
    sum_{k>=1} d_k = (D-2) S + x_0^2 + sum_{k>=1} (x_k - m)^2       (OOD.exclude_back)
    argmin_k d_k   = first largest channel (m > 0) / first smallest channel (m < 0)
    max softmax    = 1 / sum_k exp(2 m (x_k - x_ext))

The GPU parity tests (tests/test_gpu_head.py::test_lean_head_*) check the kernel itself; this file pins the algebra and
its fp32 conditioning (tight clusters: own-class distance ~0.03 against ||x||^2 ~ 9) where no GPU is available."""
import numpy as np
import pytest
import torch

from oracle import dml_oracle as O
from tests.synth import streethazards_like

f32 = np.float32


def _closed_forms(x, m, skip0):
    x = x.numpy().astype(f32)                               # [B, D, H, W]
    D = x.shape[1]
    s = [x[:, 0] * x[:, 0], np.zeros_like(x[:, 0])]         # two interleaved fp32 chains, like the kernel
    t = [np.zeros_like(x[:, 0]), np.zeros_like(x[:, 0])]
    for k in range(1, D):
        u = x[:, k] - f32(m)
        c = k % 2                                           # odd channels -> chain 1, even -> chain 0
        s[c] = (x[:, k] * x[:, k] + s[c]).astype(f32)
        t[c] = (u * u + t[c]).astype(f32)
    S, T1 = (s[0] + s[1]).astype(f32), (t[0] + t[1]).astype(f32)
    u0 = x[:, 0] - f32(m)
    if skip0 and D > 1:
        eds = (f32(D - 2) * S + (x[:, 0] * x[:, 0] + T1)).astype(f32)
    else:
        eds = (f32(D - 1) * S + (u0 * u0 + T1)).astype(f32)
    y = x if m >= 0 else -x
    label = y.argmax(axis=1)                                # first extremal channel
    ys = y[:, 1:] if (skip0 and D > 1) else y
    ext = ys.max(axis=1, keepdims=True)
    msp = 1.0 / np.exp(f32(2 * abs(m)) * (ys - ext)).astype(f32).sum(axis=1)
    return eds, label, msp.astype(f32)


@pytest.mark.parametrize("k,sigma,m", [(13, 0.7, 3.0), (13, 0.1, 3.0), (16, 0.7, 3.0), (19, 0.7, 3.0), (32, 0.7, 3.0), (3, 0.7, 3.0),
                                       (2, 0.7, 3.0), (1, 0.7, 3.0), (13, 0.7, -3.0), (13, 0.7, 0.5)])
@pytest.mark.parametrize("skip0", [False, True])
def test_closed_forms_match_the_direct_form(k, sigma, m, skip0):
    x, _ = streethazards_like(2, 48, 64, k=k, sigma=sigma, seed=k + int(10 * sigma))
    if m < 0:
        x = -x
    centers = O.make_centers(k, m)
    z = O.distance_logits(x, centers)                        # fp32, reference op order
    z64 = O.distance_logits_f64(x, centers)
    first = 1 if (skip0 and k > 1) else 0
    eds, label, msp = _closed_forms(x, m, skip0)
    np.testing.assert_allclose(eds, -(z[:, first:].sum(dim=1)).numpy(), rtol=1e-5)
    np.testing.assert_allclose(eds, -(z64[:, first:].sum(dim=1)).numpy(), rtol=2e-6)
    # label: equals the reference argmax except where the reference's own best two logits tie within rounding
    ref = z.max(dim=1)[1].numpy()
    if k > 1:
        top2 = torch.topk(z, 2, dim=1).values
        near = ((top2[:, 0] - top2[:, 1]).abs() <= 2e-6 * top2[:, 1].abs().clamp_min(1e-30)).numpy()
    else:
        near = np.zeros(ref.shape, bool)
    assert not ((label != ref) & ~near).any()
    assert np.array_equal(label, z64.max(dim=1)[1].numpy()) or ((label != z64.max(dim=1)[1].numpy()).mean() < 1e-4)
    msp64 = torch.softmax(z64[:, first:], dim=1).max(dim=1)[0].numpy()
    np.testing.assert_allclose(msp, msp64, rtol=2e-6)
