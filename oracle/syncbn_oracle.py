"""TEST INFRASTRUCTURE (CPU oracle) -- NumPy float64 restatement of the reference's synchronised batch norm:

  anomaly/lib/nn/modules/batchnorm.py:64-88    forward over the (global) batch: sum / square-sum -> mean, inv_std -> affine
  anomaly/lib/nn/modules/batchnorm.py:121-139  _compute_mean_std: inv_std = clamp(sumvar / n, eps) ** -0.5, moving averages
                                               through _tmp_running_mean / _tmp_running_var / _running_iter

and the analytic gradient of that expression (what autograd computes in the reference, including the clamp's zero
gradient).  Pinned by tests/test_oracle_syncbn.py against tests/golden/syncbn.npz (the reference module itself, two
training steps).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module."""
import numpy as np


class SyncBNState:
    """the buffers of _SynchronizedBatchNorm.__init__ (batchnorm.py:39-55)"""

    def __init__(self, C, momentum=0.001):
        self.keep = 1.0 - momentum
        self.running_mean = np.zeros(C)
        self.running_var = np.ones(C)
        self.running_iter = 1.0
        self.tmp_running_mean = self.running_mean * self.running_iter
        self.tmp_running_var = self.running_var * self.running_iter


def forward(x, weight, bias, eps=1e-5, state=None):
    """x [B, C, ...] (the batch of ALL devices) -> (y, cache)"""
    x = np.asarray(x, np.float64)
    C = x.shape[1]
    xr = x.reshape(x.shape[0], C, -1)
    n = xr.shape[0] * xr.shape[2]
    assert n > 1
    s = xr.sum(axis=(0, 2))
    ss = (xr ** 2).sum(axis=(0, 2))
    mean = s / n
    sumvar = ss - s * mean
    unbias_var = sumvar / (n - 1)
    bias_var = sumvar / n
    clamped = bias_var < eps
    inv_std = np.maximum(bias_var, eps) ** -0.5
    if state is not None:
        state.tmp_running_mean = state.tmp_running_mean * state.keep + mean
        state.tmp_running_var = state.tmp_running_var * state.keep + unbias_var
        state.running_iter = state.running_iter * state.keep + 1
        state.running_mean = state.tmp_running_mean / state.running_iter
        state.running_var = state.tmp_running_var / state.running_iter
    w = np.ones(C) if weight is None else np.asarray(weight, np.float64)
    b = np.zeros(C) if bias is None else np.asarray(bias, np.float64)
    y = (xr - mean[None, :, None]) * (inv_std * w)[None, :, None] + b[None, :, None]
    return y.reshape(x.shape), (xr, mean, inv_std, clamped, w, n)


def backward(dy, cache):
    """(dx, dweight, dbias) of sum(y * dy) for the global batch"""
    xr, mean, inv_std, clamped, w, n = cache
    d = np.asarray(dy, np.float64).reshape(xr.shape)
    xc = xr - mean[None, :, None]
    s0 = d.sum(axis=(0, 2))
    s1 = (d * xc).sum(axis=(0, 2))
    k1 = np.where(clamped, 0.0, s1 / n * inv_std ** 2)
    dx = (w * inv_std)[None, :, None] * (d - (s0 / n)[None, :, None] - xc * k1[None, :, None])
    return dx.reshape(np.asarray(dy).shape), s1 * inv_std, s0
