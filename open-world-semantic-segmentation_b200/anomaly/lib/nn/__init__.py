"""``lib.nn`` of the reference (anomaly/lib/nn/__init__.py) for DistributedDataParallel: synchronised batch norm whose
statistics are all-reduced over NCCL instead of being piped between the replicas of one DataParallel process."""
from .batchnorm import (SynchronizedBatchNorm1d, SynchronizedBatchNorm2d, SynchronizedBatchNorm3d,  # noqa: F401
                        patch_replication_callback, convert_model)
