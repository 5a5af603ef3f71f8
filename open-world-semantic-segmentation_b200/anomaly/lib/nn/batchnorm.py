"""Synchronised batch normalisation, one process per GPU.

Reference: anomaly/lib/nn/modules/batchnorm.py:38-139 (``_SynchronizedBatchNorm``), used as ``BatchNorm2d`` by the
PSPNet / ResNet modules of anomaly/models.  There the replicas of one ``DataParallel`` process send (sum, square-sum,
size) to the master replica through queues; here every rank computes its partial sums with ``dml_bn_stats``, one NCCL
all-reduce of 2C + 1 doubles makes them global, and ``dml_bn_finalize`` / ``dml_bn_apply`` finish identically on every
rank -- same formulas (``inv_std = clamp(var, eps) ** -0.5``, moving averages through ``_tmp_running_*`` /
``_running_iter``), same constructor defaults (momentum 0.001).  The backward pass all-reduces (sum dy, sum dy (x - mean))
the same way.  As in the reference, evaluation mode and non-parallel training (world size 1, unless ``always_sync``) call
``F.batch_norm``.  Weight / bias gradients are the rank's own sums: DistributedDataParallel reduces them with the
other parameter gradients."""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch.nn.modules.batchnorm import _BatchNorm

from ...._lib import check, lib, ptr, require_cuda, stream_ptr

__all__ = ["SynchronizedBatchNorm1d", "SynchronizedBatchNorm2d", "SynchronizedBatchNorm3d", "patch_replication_callback",
           "convert_model"]


def _world(group):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class _SyncBNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, module, group):
        require_cuda(x, "input")
        if x.dtype != torch.float32:
            raise ValueError("SynchronizedBatchNorm: float32 input expected")
        x = x.contiguous()
        B, Cn = x.shape[0], x.shape[1]
        HW = x.numel() // max(B * Cn, 1)
        dev = x.device
        sums = torch.empty(2 * Cn + 1, dtype=torch.float64, device=dev)
        ws = torch.empty(lib().dml_bn_workspace_bytes(B, Cn, HW), dtype=torch.uint8, device=dev)
        mean = torch.empty(Cn, dtype=torch.float32, device=dev)
        inv_std = torch.empty(Cn, dtype=torch.float32, device=dev)
        clamped = torch.empty(Cn, dtype=torch.uint8, device=dev)
        y = torch.empty_like(x)
        w = weight.detach().contiguous() if weight is not None else None
        b = bias.detach().contiguous() if bias is not None else None
        track = module.training and module.track_running_stats
        with torch.cuda.device(dev):
            s = stream_ptr(dev)
            check(lib().dml_bn_stats(ptr(x), None, None, B, Cn, HW, ptr(sums), ptr(ws), ws.numel(), s), "dml_bn_stats")
            sums[2 * Cn:].fill_(float(B * HW))
            if _world(group) > 1:
                dist.all_reduce(sums, group=group)
            check(lib().dml_bn_finalize(ptr(sums), ptr(sums[2 * Cn:]), module.eps, module.momentum, Cn,
                                        ptr(module._tmp_running_mean) if track else None,
                                        ptr(module._tmp_running_var) if track else None,
                                        ptr(module._running_iter) if track else None,
                                        ptr(module.running_mean) if track else None, ptr(module.running_var) if track else None,
                                        ptr(mean), ptr(inv_std), ptr(clamped), s), "dml_bn_finalize")
            check(lib().dml_bn_apply(ptr(x), ptr(mean), ptr(inv_std), ptr(w), ptr(b), B, Cn, HW, ptr(y), s), "dml_bn_apply")
        ctx.save_for_backward(x, w, mean, inv_std, clamped, sums)
        ctx.group = group
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, mean, inv_std, clamped, sums = ctx.saved_tensors
        dy = dy.contiguous()
        B, Cn = x.shape[0], x.shape[1]
        HW = x.numel() // max(B * Cn, 1)
        dev = x.device
        bsums = torch.empty(2 * Cn, dtype=torch.float64, device=dev)
        ws = torch.empty(lib().dml_bn_workspace_bytes(B, Cn, HW), dtype=torch.uint8, device=dev)
        dx = torch.empty_like(x)
        with torch.cuda.device(dev):
            s = stream_ptr(dev)
            check(lib().dml_bn_stats(ptr(x), ptr(dy), ptr(mean), B, Cn, HW, ptr(bsums), ptr(ws), ws.numel(), s), "dml_bn_stats")
            # parameter gradients from this rank's own pixels (DDP reduces them): d/dw = inv_std sum dy (x - mean), d/db = sum dy
            dw = (bsums[Cn:] * inv_std.double()).float() if w is not None else None
            db = bsums[:Cn].float() if ctx.has_bias else None
            if _world(ctx.group) > 1:
                dist.all_reduce(bsums, group=ctx.group)
            check(lib().dml_bn_backward_apply(ptr(x), ptr(dy), ptr(mean), ptr(inv_std), ptr(w), ptr(clamped), ptr(bsums),
                                              ptr(sums[2 * Cn:]), B, Cn, HW, ptr(dx), s), "dml_bn_backward_apply")
        return dx, dw, db, None, None


class _SynchronizedBatchNorm(_BatchNorm):
    """anomaly/lib/nn/modules/batchnorm.py:38-139.  ``process_group``: the ranks that share statistics (default: all);
    ``always_sync``: run the synchronised kernels even with a single rank (testing / single-GPU training)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.001, affine=True, process_group=None, always_sync=False):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine)
        self.process_group = process_group
        self.always_sync = always_sync
        self._moving_average_fraction = 1. - momentum
        self.register_buffer("_tmp_running_mean", torch.zeros(self.num_features))
        self.register_buffer("_tmp_running_var", torch.ones(self.num_features))
        self.register_buffer("_running_iter", torch.ones(1))
        self._tmp_running_mean = self.running_mean.clone() * self._running_iter
        self._tmp_running_var = self.running_var.clone() * self._running_iter

    @property
    def _is_parallel(self):
        return self.always_sync or _world(self.process_group) > 1

    def forward(self, input):
        # batchnorm.py:58-62: evaluation mode or non-parallel computation -> PyTorch's implementation
        if not (self._is_parallel and self.training):
            return F.batch_norm(input, self.running_mean, self.running_var, self.weight, self.bias, self.training, self.momentum,
                                self.eps)
        self._check_input_dim(input)
        return _SyncBNFunction.apply(input, self.weight, self.bias, self, self.process_group)


class SynchronizedBatchNorm1d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 2 and input.dim() != 3:
            raise ValueError("expected 2D or 3D input (got {}D input)".format(input.dim()))


class SynchronizedBatchNorm2d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError("expected 4D input (got {}D input)".format(input.dim()))


class SynchronizedBatchNorm3d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 5:
            raise ValueError("expected 5D input (got {}D input)".format(input.dim()))


def patch_replication_callback(data_parallel):
    """anomaly/lib/nn/modules/replicate.py: hooks the master / slave pipes into ``DataParallel.replicate``.  With one
    process per GPU there are no replicas to wire up: kept so that training scripts import and call it unchanged."""
    return data_parallel


def convert_model(module, process_group=None):
    """nn.BatchNorm{1,2,3}d -> SynchronizedBatchNorm{1,2,3}d (parameters and running statistics carried over), in place of
    the reference's build-time ``BatchNorm2d = SynchronizedBatchNorm2d`` aliasing (anomaly/models/resnet.py:8)."""
    mapping = {torch.nn.BatchNorm1d: SynchronizedBatchNorm1d, torch.nn.BatchNorm2d: SynchronizedBatchNorm2d,
               torch.nn.BatchNorm3d: SynchronizedBatchNorm3d}
    for src, dst in mapping.items():
        if type(module) is src:
            new = dst(module.num_features, module.eps, module.momentum if module.momentum is not None else 0.001, module.affine,
                      process_group=process_group)
            if module.affine:
                new.weight, new.bias = module.weight, module.bias
            new.running_mean.copy_(module.running_mean)
            new.running_var.copy_(module.running_var)
            new._tmp_running_mean = new.running_mean.clone() * new._running_iter
            new._tmp_running_var = new.running_var.clone() * new._running_iter
            return new.to(module.running_mean.device)
    for name, child in module.named_children():
        module.add_module(name, convert_model(child, process_group))
    return module
