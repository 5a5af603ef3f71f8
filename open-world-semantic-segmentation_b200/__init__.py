"""dml_b200 -- B200-native (sm_100a) implementation of DMLNet's per-pixel metric-learning hot path.

Kernels live in ``csrc/`` behind the C ABI of ``include/dml_b200.h`` (``libdml_b200.so``);
this package is the host-side mirror of the reference's Python operator surface:

* ``dml_b200.head``      fused distance + score head (labels, EDS, MMSP, confusion, NPM)
* ``dml_b200.ood``       exact AUROC / AUPR / FPR@95 (radix sort + tie-aware scan)
* ``dml_b200.anomaly``   drop-ins for ``anomaly/{anom_utils,utils,models}.py`` + the score block
* ``dml_b200.deeplab``   drop-ins for ``DeepLabV3Plus-Pytorch/{network,utils/loss,metrics}``

There is no CPU / PyTorch fallback: every entry point raises if the CUDA library is missing.
"""
from . import _lib  # noqa: F401
from ._lib import DmlError, load_library  # noqa: F401
from . import head, ood  # noqa: F401
from . import autograd, anomaly, deeplab, prototypes  # noqa: F401
from .autograd import distance_logits, dml_loss  # noqa: F401
from .head import (HeadOutput, confusion_counts, dml_head, dml_multiscale_head, finalize_scores,  # noqa: F401
                   multiscale_average, plm_merge)

__version__ = "0.1.0"
