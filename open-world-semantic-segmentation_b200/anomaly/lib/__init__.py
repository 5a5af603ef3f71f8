"""Mirror of the reference's ``anomaly/lib`` package for the one-process-per-GPU (DDP / NCCL) training set-up."""
from . import nn  # noqa: F401
