"""Drop-in surface of the reference's ``anomaly/`` sub-project for the DML hot path
(``anom_utils``, the embedding decoder head, the score block of ``evaluate`` and the
accuracy / IoU counters), backed by the CUDA kernels of libdml_b200.so."""
from . import anom_utils, dataset, eval_ood, models, utils  # noqa: F401
