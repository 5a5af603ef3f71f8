"""ctypes binding of libdml_b200.so (the C ABI in include/dml_b200.h).

The product path has NO fallback: if the shared library is missing or a kernel launch
fails, this module raises.  PyTorch is only used for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdml_b200.so")

DML_OK = 0
ABI_VERSION = 4

c_f32p = C.c_void_p
c_void_p = C.c_void_p


class HeadParams(C.Structure):
    """Mirror of ``dml_head_params`` (include/dml_b200.h)."""
    _fields_ = [
        ("struct_bytes", C.c_uint32),
        ("B", C.c_int32), ("D", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("x", C.c_void_p), ("mu", C.c_void_p),
        ("diag_m", C.c_float), ("input_is_logits", C.c_int32), ("score_first_class", C.c_int32), ("eds_clamp", C.c_float),
        ("mu_novel", C.c_void_p), ("n_novel", C.c_int32), ("novel_label_base", C.c_int32),
        ("novel_thr", C.c_double),
        ("logits", C.c_void_p), ("label_u8", C.c_void_p), ("label_i64", C.c_void_p),
        ("maxlogit", C.c_void_p), ("eds", C.c_void_p), ("msp", C.c_void_p),
        ("features_nhwc", C.c_void_p), ("novel_dist", C.c_void_p), ("minmax", C.c_void_p),
        ("want_eds_minmax", C.c_int32), ("want_msp_minmax", C.c_int32),
        ("gt_u8", C.c_void_p), ("gt_i64", C.c_void_p), ("confusion", C.c_void_p),
        ("conf_rows", C.c_int32), ("conf_cols", C.c_int32), ("reference_order", C.c_int32),
    ]


MAX_SCALES = 8


class MultiscaleParams(C.Structure):
    """Mirror of ``dml_multiscale_params`` (include/dml_b200.h)."""
    _fields_ = [
        ("struct_bytes", C.c_uint32),
        ("B", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("n_scales", C.c_int32),
        ("z", C.c_void_p * MAX_SCALES), ("h", C.c_int32 * MAX_SCALES), ("w", C.c_int32 * MAX_SCALES),
        ("reciprocal_average", C.c_int32), ("score_first_class", C.c_int32), ("eds_clamp", C.c_float),
        ("scores", C.c_void_p), ("label_u8", C.c_void_p), ("label_i64", C.c_void_p),
        ("maxlogit", C.c_void_p), ("eds", C.c_void_p), ("msp", C.c_void_p), ("minmax", C.c_void_p),
        ("want_eds_minmax", C.c_int32), ("want_msp_minmax", C.c_int32),
        ("gt_u8", C.c_void_p), ("gt_i64", C.c_void_p), ("confusion", C.c_void_p),
        ("conf_rows", C.c_int32), ("conf_cols", C.c_int32),
    ]


class OodResult(C.Structure):
    """Mirror of ``dml_ood_result``."""
    _fields_ = [("auroc", C.c_double), ("aupr", C.c_double), ("fpr", C.c_double),
                ("n_pos", C.c_longlong), ("n_neg", C.c_longlong), ("n_nan", C.c_longlong),
                ("n_groups", C.c_longlong)]


OOD_RESULT_WORDS = C.sizeof(OodResult) // 8  # 7 x 8 bytes

# name -> (restype, argtypes); must list every symbol include/dml_b200.h declares
SIGNATURES = {
    "dml_abi_version": (C.c_int, []),
    "dml_error_string": (C.c_char_p, [C.c_int]),
    "dml_last_cuda_error": (C.c_int, []),
    "dml_max_dim": (C.c_int, []),
    "dml_kernel_launches": (C.c_ulonglong, []),
    "dml_head_forward": (C.c_int, [C.POINTER(HeadParams), C.c_void_p]),
    "dml_multiscale_head_forward": (C.c_int, [C.POINTER(MultiscaleParams), C.c_void_p]),
    "dml_conv1x1_head_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dml_scores_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_float,
                                      C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dml_confusion": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                C.c_void_p, C.c_void_p]),
    "dml_plm_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "dml_loss_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "dml_loss_forward": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "dml_loss_backward": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "dml_class_sums_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int64]),
    "dml_class_sums": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64,
                                 C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dml_ood_keystats": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "dml_ood_keygen": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                 C.c_void_p, C.c_int32, C.c_uint32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_size_t,
                                 C.c_void_p]),
    "dml_ood_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int64]),
    "dml_ood_eval_segments": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_void_p,
                                        C.c_size_t, C.c_int32, C.c_void_p, C.c_void_p]),
    "dml_ood_pool_histograms": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_size_t,
                                          C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "dml_ood_roc_fpr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_void_p, C.c_size_t,
                                  C.c_void_p, C.c_void_p]),
    "dml_ood_sort": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t,
                               C.POINTER(C.c_void_p), C.c_void_p]),
    "dml_ood_partition_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "dml_ood_partition": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_size_t, C.c_void_p]),
    "dml_ood_lower_bound": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "dml_ood_count_positive": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "dml_ood_rank_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "dml_ood_rank_segments": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                        C.c_void_p, C.c_int32, C.c_uint32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int32, C.c_double,
                                        C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "dml_ood_rank_export_positives": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                                                C.c_void_p]),
    "dml_ood_pos_compact": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "dml_bn_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64]),
    "dml_bn_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t,
                               C.c_void_p]),
    "dml_bn_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dml_bn_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p,
                               C.c_void_p]),
    "dml_bn_backward_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "dml_resize_ksize": (C.c_int32, [C.c_int32, C.c_int32]),
    "dml_resize_coeffs": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    "dml_resize_bilinear_normalize": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                                C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "dml_ood_slots_gather": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "dml_ood_unique_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "dml_ood_unique_counts": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_void_p]),
    "dml_ood_bucket_rank_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "dml_ood_bucket_rank": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p,
                                      C.c_size_t, C.c_void_p]),
    "dml_ood_pooled_scan_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "dml_ood_pooled_scan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_double,
                                      C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "dml_ood_scan_range": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_void_p, C.c_size_t,
                                     C.c_void_p, C.c_void_p]),
}

_lib = None


class DmlError(RuntimeError):
    pass


def load_library(path: str | None = None):
    """dlopen libdml_b200.so and bind every entry point.  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("DML_B200_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise DmlError(
            f"{path} not found: build it with `python open-world-semantic-segmentation_b200/build.py` "
            "(or __graft_entry__.build()).  There is no CPU/PyTorch fallback for the DML hot path.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.dml_abi_version() != ABI_VERSION:
        raise DmlError(f"ABI mismatch: library {lib.dml_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def lib():
    return load_library()


def check(rc: int, what: str = ""):
    if rc != DML_OK:
        l = lib()
        msg = l.dml_error_string(rc).decode()
        if rc == -3:
            msg += f" (cudaError {l.dml_last_cuda_error()})"
        raise DmlError(f"{what or 'dml call'} failed: {msg} [{rc}]")


def ptr(t):
    """Device pointer of a tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, name: str = "tensor"):
    if not t.is_cuda:
        raise DmlError(f"{name} must be a CUDA tensor: the DML hot path runs only on the GPU (no CPU fallback)")
