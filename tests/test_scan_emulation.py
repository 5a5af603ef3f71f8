"""CPU tier: the per-thread code of the GPU group scan (csrc/ood_scan_thread.cuh is __host__ __device__) is compiled
with g++ and run thread by thread / tile by tile by tests/host/scan_emulation.cpp.  For every thread the bit-mask
form must equal the one-key-at-a-time state machine exactly, and the assembled (auroc, aupr, fpr) must equal the
oracle (anomaly/anom_utils.py:67-78 semantics).  Covers ties, plateaus, ragged tails, ranges with carried counts
(the multi-GPU scan_range form) and the degenerate single-class cases."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import dml_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("scan_emul") / "libscan_emul.so"
    src = os.path.join(ROOT, "tests", "host", "scan_emulation.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(out), src], check=True)
    lib = ctypes.CDLL(str(out))
    lib.scan_emulate.restype = ctypes.c_longlong
    lib.scan_emulate.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong,
                                 ctypes.c_longlong, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
    lib.pack_keys.restype = None
    lib.pack_keys.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_uint32,
                              ctypes.c_void_p, ctypes.c_void_p]
    return lib


def _pack(emu, values, pos, kind, key_base):
    values = np.ascontiguousarray(values, dtype=np.float32)
    pos = np.ascontiguousarray(pos, dtype=np.uint8)
    keys = np.zeros(values.size, np.uint32)
    counts = np.zeros(2, np.int64)
    emu.pack_keys(values.ctypes.data, pos.ctypes.data, values.size, kind, key_base, keys.ctypes.data, counts.ctypes.data)
    return keys, counts


def _keys(conf, pos):
    """packed ranking keys of dml_ood_keygen (score_kind 0, key_base = sortable(+0.0)), sorted ascending"""
    c = np.where(conf == 0, np.float32(0), conf.astype(np.float32))
    bits = c.view(np.uint32)
    srt = np.where(bits & 0x80000000, ~bits, bits | np.uint32(0x80000000)).astype(np.uint32)
    rel = (srt.astype(np.int64) - 0x80000000)
    assert (rel >= 0).all() and (rel < (1 << 31)).all()
    return np.sort(((rel.astype(np.uint64) << 1) | pos.astype(np.uint64)).astype(np.uint32))


def _run(emu, keys, pos_before, idx_before, total_pos, total_n, level=0.95):
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    out = np.zeros(3, np.float64)
    part = np.zeros(10, np.int64)
    bad = emu.scan_emulate(keys.ctypes.data, keys.size, pos_before, idx_before, total_pos, total_n, level,
                           out.ctypes.data, part.ctypes.data)
    return bad, out, part


def _case(seed, n, kind):
    rng = np.random.default_rng(seed)
    conf = rng.random(n).astype(np.float32)
    pos = rng.random(n) < (0.01 if kind != "dense" else 0.5)
    if kind == "ties":
        conf = (np.round(conf * 12) / 12).astype(np.float32)
    elif kind == "plateau":
        conf[rng.random(n) < 0.3] = 1.0
        conf[rng.random(n) < 0.05] = 0.0
    elif kind == "separated":                       # positives at low conf (the well-separated OOD case)
        conf[pos] *= np.float32(0.05)
    elif kind == "pairs":                           # every score twice: one positive + one negative per group
        conf = np.repeat(rng.random((n + 1) // 2).astype(np.float32), 2)[:n]
        pos = (np.arange(n) % 2) == 0
    elif kind == "allsame":
        conf[:] = 0.25
    return conf, pos


@pytest.mark.parametrize("kind", ["plain", "ties", "plateau", "separated", "dense", "pairs", "allsame"])
@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 255, 4095, 4096, 4097, 8192 + 5, 70001])
def test_scan_emulation_matches_oracle(emu, kind, n):
    conf, pos = _case(n * 31 + len(kind), n, kind)
    keys = _keys(conf, pos)
    P = int(pos.sum())
    bad, out, _ = _run(emu, keys, 0, 0, P, n)
    assert bad == 0, "bit-mask form differs from the one-key-at-a-time form"
    gt = np.where(pos, 13, 3).astype(np.int64)
    ref = O.eval_ood_measure(conf, gt, (13,))
    if ref is None:
        assert np.isnan(out).all()
    else:
        np.testing.assert_allclose(out, ref, rtol=0, atol=1e-12)


@pytest.mark.parametrize("level", [0.5, 0.9, 0.95, 1.0])
def test_scan_emulation_recall_levels_and_few_positives(emu, level):
    for seed, n, npos in [(1, 5000, 1), (2, 5000, 2), (3, 5000, 20), (4, 33, 33 - 1), (5, 4096 * 3, 19)]:
        rng = np.random.default_rng(seed)
        conf = (np.round(rng.random(n) * 200) / 200).astype(np.float32)
        pos = np.zeros(n, bool)
        pos[rng.choice(n, npos, replace=False)] = True
        bad, out, _ = _run(emu, _keys(conf, pos), 0, 0, npos, n, level)
        assert bad == 0
        labels = pos.astype(np.int32)
        fpr = O.fpr_and_fdr_at_recall(labels, -conf, level)
        assert out[2] == pytest.approx(fpr, abs=1e-12)


@pytest.mark.parametrize("kind", ["plain", "ties", "plateau", "separated"])
def test_scan_emulation_ranges_combine_like_one_segment(emu, kind):
    """dml_ood_scan_range form: the ranking cut at group boundaries into 3 ranges with carried counts; the partials
    combine (ood.combine_partials) to the single-segment result bit for bit."""
    from dml_b200 import ood
    n = 30011
    conf, pos = _case(77, n, kind)
    keys = _keys(conf, pos)
    P = int(pos.sum())
    bad, whole, _ = _run(emu, keys, 0, 0, P, n)
    assert bad == 0
    cuts = [0]
    for c in (n // 3, 2 * n // 3):
        while c < n and (keys[c] >> 1) == (keys[c - 1] >> 1):   # move the cut to a group boundary
            c += 1
        cuts.append(c)
    cuts.append(n)
    partials = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        bad, _, part = _run(emu, keys[a:b], int((keys[:a] & 1).sum()), a, P, n)
        assert bad == 0
        partials.append(part)
    au, ap, fp, groups = ood.combine_partials(partials, P, n)
    assert (au, fp) == (whole[0], whole[2])
    assert ap == pytest.approx(whole[1], abs=1e-14)
    assert groups == np.unique(keys >> 1).size


def test_pack_key_is_order_preserving_and_matches_the_numpy_double(emu):
    """the kernels' pack_key (csrc/ood_scan_thread.cuh) on the CPU: ascending key order == ascending conf (kind 0) /
    descending score (kind 1), equal keys <=> equal values with -0 == +0, positives after negatives inside a value,
    NaNs and window violations counted; and it is the packing tests/np_ops.py and `_keys` above assume."""
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.random(5000), [0.0, -0.0, 1.0, 1.0, 2.0 ** -149, 2.0 ** -126, 0.5, 0.5]]).astype(np.float32)
    pos = (rng.random(v.size) < 0.3)
    keys, counts = _pack(emu, v, pos, 0, 0x80000000)
    assert counts.tolist() == [0, 0]
    np.testing.assert_array_equal(np.sort(keys), _keys(v, pos))
    order = np.argsort(keys, kind="stable")
    assert (np.diff(v[order].astype(np.float64)) >= 0).all()                 # ascending conf
    same = (keys[order][1:] >> 1) == (keys[order][:-1] >> 1)
    np.testing.assert_array_equal(same, v[order][1:] == v[order][:-1])      # groups == equal values (+-0 merged)
    assert ((keys[order][1:] & 1)[same] >= (keys[order][:-1] & 1)[same]).all()   # negatives first inside a group
    from tests.np_ops import NumpyOps
    import torch
    k2, st = NumpyOps().make_keys(torch.from_numpy(v), torch.from_numpy(np.where(pos, 13, 2)), (13,), 0x80000000)
    np.testing.assert_array_equal(keys, k2.numpy().view(np.uint32))
    assert int(st[0]) == int(pos.sum())
    # plain scores of either sign (kind 1): descending score order, window chosen from the data like measures_from_scores
    sc = (np.abs(rng.standard_normal(4000)) * 3 + 0.01).astype(np.float32)     # same sign: the keys span < 2^31 values
    f = -sc
    bits = np.where(f == 0, np.float32(0), f).view(np.uint32)
    srt = np.where(bits & 0x80000000, ~bits, bits | np.uint32(0x80000000)).astype(np.uint32)
    base = int(srt.min())
    k1, c1 = _pack(emu, sc, np.zeros(sc.size, bool), 1, base)
    assert c1.tolist() == [0, 0]
    o1 = np.argsort(k1, kind="stable")
    assert (np.diff(sc[o1].astype(np.float64)) <= 0).all()
    # mixed-sign scores span more than 2^31 sortable values: counted, never silently wrapped (measures_from_scores then
    # ranks the two signs as separate ranges)
    wide = np.float32([-3.0, 2.0, 1e-3, -1e-3])
    wb = np.where(-wide == 0, np.float32(0), -wide).view(np.uint32)
    wsrt = np.where(wb & 0x80000000, ~wb, wb | np.uint32(0x80000000)).astype(np.uint32)
    _, cw = _pack(emu, wide, [0, 0, 0, 0], 1, int(wsrt.min()))
    assert cw[1] > 0
    # NaN: counted, ranked last; out of window: counted and clamped to the window's ends
    kn, cn = _pack(emu, np.float32([0.25, np.nan, -1.0, 0.5]), [0, 1, 0, 0], 0, 0x80000000)
    assert cn.tolist() == [1, 1]
    assert kn[1] == 0xFFFFFFFF and kn[2] == 0 and kn[0] < kn[3] < kn[1]

