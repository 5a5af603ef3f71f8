"""CPU tier: the scalar arithmetic of the fused multi-scale head (csrc/dml_bilinear.cuh, __host__ __device__ -- the
functions head_kernel<K, HEAD_MS> calls per tap) compiled with g++ and run pixel by pixel by tests/host/ms_emulation.cpp
must reproduce, BIT FOR BIT, a torch-CPU replay of the reference loop
    scores += F.interpolate(z_s, segSize, mode='bilinear', align_corners=False) / n_scales
(anomaly/models/models.py:659-661 + anomaly/eval_ood_traditional.py:198-208) -- up- and down-sampling scales, odd sizes,
edge clamping, both averaging conventions.  The GPU tests (tests/test_gpu_multiscale.py) check the kernel itself."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import dml_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("ms_emul") / "libms_emul.so"
    src = os.path.join(ROOT, "tests", "host", "ms_emulation.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(out), src], check=True)
    lib = ctypes.CDLL(str(out))
    lib.ms_emulate.restype = None
    return lib


def _emulate(emu, zs, H, W, reciprocal=False):
    B, K = zs[0].shape[:2]
    arrs = [np.ascontiguousarray(z.numpy(), dtype=np.float32) for z in zs]
    n = len(arrs)
    ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrs])
    hs = (ctypes.c_int * n)(*[a.shape[2] for a in arrs])
    ws = (ctypes.c_int * n)(*[a.shape[3] for a in arrs])
    out = np.zeros((B, K, H, W), np.float32)
    emu.ms_emulate(ptrs, hs, ws, n, B, K, H, W, 1 if reciprocal else 0, out.ctypes.data_as(ctypes.c_void_p))
    return out


CASES = [
    # (H, W, low-resolution sizes).  Outputs are kept above ~10^5 elements: for small outputs torch's CPU kernel itself takes
    # a differently contracted (serial) path whose results differ from its own large-output path -- and from torch CUDA -- in
    # the last ulp (see test_small_outputs_...); every shape of the reference's evaluation is far above that size.
    (96, 160, [(5, 9), (7, 11), (8, 13), (9, 15), (10, 17)]),                  # the golden fixture's shape (stride-8 maps)
    (180, 320, [(38, 67), (47, 84), (57, 100), (66, 117), (71, 125)]),         # StreetHazards' five scales, reduced target
    (137, 253, [(5, 7), (137, 253), (280, 511)]),                              # up-, same- and down-sampling, odd sizes
    (128, 136, [(1, 5), (2, 3), (3, 1)]),                                      # degenerate maps: the taps clamp at the edges
    (133, 265, [(11, 21)]),                                                    # single scale: division by 1
]


@pytest.mark.parametrize("H,W,sizes", CASES)
@pytest.mark.parametrize("K", [13, 3])
def test_multiscale_arithmetic_is_bit_identical_to_the_torch_cpu_replay(emu, H, W, sizes, K):
    g = torch.Generator().manual_seed(H * 1000 + W + K)
    zs = [torch.randn(2, K, h, w, generator=g) * 30 - 60 for h, w in sizes]       # logits of the scale of -distances
    n = len(zs)
    ref = torch.zeros(2, K, H, W)
    for z in zs:
        ref = ref + F.interpolate(z, size=(H, W), mode="bilinear", align_corners=False) / n
    got = _emulate(emu, zs, H, W)
    assert np.array_equal(got, ref.numpy()), f"{(got != ref.numpy()).mean():.3g} of the values differ"
    # torch's CUDA division by a scalar multiplies by fl(1/n) (`reciprocal_average`): replayed on the CPU as a product
    ref_r = torch.zeros(2, K, H, W)
    for z in zs:
        ref_r = ref_r + F.interpolate(z, size=(H, W), mode="bilinear", align_corners=False) * np.float32(1.0 / n)
    assert np.array_equal(_emulate(emu, zs, H, W, reciprocal=True), ref_r.numpy())


def test_multiscale_emulation_matches_the_reference_golden(emu, golden):
    """end of the chain: stride-8 embeddings captured from the reference's evaluate() -> oracle distance logits ->
    emulated kernel arithmetic == the oracle's (torch-CPU) multi-scale scores, hence the reference's pred."""
    g = golden("evaluate_anomaly.npz")
    c = O.make_centers(13)
    for i in range(2):
        lows = [torch.from_numpy(g[f"img{i}_low{s}"]) for s in range(5)]
        seg = g[f"img{i}_seg"]
        zs = [O.distance_logits(x, c) for x in lows]
        got = _emulate(emu, zs, *seg.shape)
        scores, _ = O.multiscale_scores(lows, c, seg.shape)
        assert np.array_equal(got, scores.numpy())
        np.testing.assert_array_equal(got.argmax(axis=1)[0], g[f"img{i}_pred"])


def test_multiscale_emulation_config0_full_shape(emu, golden):
    """BASELINE.json configs[0] at its real shape (720x1280, the five StreetHazards scales): the kernel's scalar arithmetic
    applied to the stride-8 embeddings the reference produced gives exactly the reference's pred and conf map."""
    g = golden("config0_full_shape.npz")
    c = O.make_centers(13)
    zs = [O.distance_logits(torch.from_numpy(g[f"img0_low{s}"]), c) for s in range(5)]
    got = _emulate(emu, zs, 720, 1280)
    np.testing.assert_array_equal(got.argmax(axis=1)[0], g["img0_pred"])
    conf = O.score_dissum(torch.from_numpy(got), 400.0)
    np.testing.assert_array_equal(conf[::8, ::8], g["img0_conf_sub8"])
    assert float(conf.astype(np.float64).sum()) == float(g["img0_conf_sum"])


def test_small_outputs_agree_to_the_last_ulps(emu):
    """below its parallel grain size torch's CPU upsample runs a serial loop with a different FMA contraction (its result
    then differs from its own large-output path, which is the one the kernel arithmetic reproduces): closeness only."""
    g = torch.Generator().manual_seed(5)
    z = torch.randn(1, 13, 9, 15, generator=g) * 30 - 60
    ref = F.interpolate(z, size=(37, 53), mode="bilinear", align_corners=False)
    np.testing.assert_allclose(_emulate(emu, [z], 37, 53), ref.numpy(), rtol=0, atol=1e-4)   # values ~ -60 +- 30: a few ulps

