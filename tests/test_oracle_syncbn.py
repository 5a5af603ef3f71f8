"""CPU: the SyncBN oracle against the reference module's own outputs (tests/golden/syncbn.npz: two training steps of
anomaly/lib/nn/modules/batchnorm.py on its parallel branch, autograd gradients, moving averages)."""
import os

import numpy as np

from oracle import syncbn_oracle as S

GOLD = os.path.join(os.path.dirname(__file__), "golden", "syncbn.npz")


def test_oracle_reproduces_the_reference_module():
    g = np.load(GOLD)
    C = g["weight"].size
    st = S.SyncBNState(C, float(g["momentum"]))
    for step in range(2):
        y, cache = S.forward(g[f"x{step}"], g["weight"], g["bias"], float(g["eps"]), st)
        np.testing.assert_allclose(y, g[f"y{step}"], rtol=2e-5, atol=2e-5)
        dx, dw, db = S.backward(g[f"g{step}"], cache)
        scale = np.abs(g[f"dx{step}"]).max(axis=(0, 2, 3), keepdims=True) + 1e-30
        assert (np.abs(dx - g[f"dx{step}"]) / scale).max() < 2e-5          # relative to the channel's largest gradient
        np.testing.assert_allclose(dw, g[f"dw{step}"], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(db, g[f"db{step}"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(st.running_mean, g[f"running_mean{step}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(st.running_var, g[f"running_var{step}"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(st.running_iter, g[f"running_iter{step}"][0], rtol=1e-6)
    # the constant channel: clamp active, gradient has no variance term
    _, cache = S.forward(g["x0"], g["weight"], g["bias"], float(g["eps"]))
    assert cache[3][5] and not cache[3][:5].any()
