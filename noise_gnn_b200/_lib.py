"""ctypes binding of libngnn_b200.so — the C ABI declared in include/ngnn_b200.h.

There is no CPU fallback: if the library cannot be found or built, or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p
from pathlib import Path

from . import _build

_P = c_void_p

# name -> (restype, argtypes); mirrors include/ngnn_b200.h one to one
SIGNATURES = {
    "ngnn_version": (c_int32, []),
    "ngnn_last_error": (c_int32, [c_char_p, c_size_t]),
    "ngnn_device_supported": (c_int32, []),
    "ngnn_launch_count": (c_uint64, []),
    "ngnn_coo_to_csr_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "ngnn_coo_to_csr": (c_int32, [_P, _P, c_int64, c_int64, _P, _P, _P, _P, c_size_t, _P]),
    "ngnn_csr_transpose_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "ngnn_csr_transpose": (c_int32, [_P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, c_size_t, _P]),
    "ngnn_csr_to_coo": (c_int32, [_P, _P, c_int64, c_int64, _P, _P]),
    "ngnn_gather_rows": (c_int32, [_P, c_int64, _P, c_int64, c_int64, _P, c_int64, _P]),
    "ngnn_sage_agg_fwd": (c_int32, [_P, _P, _P, c_int64, c_int64, c_int64, _P, c_int64, _P, _P, c_int64, _P]),
    "ngnn_set_tuning": (c_int32, [c_int32, c_int32]),
    "ngnn_gcn_agg_fwd": (c_int32, [_P, _P, _P, c_int64, c_int64, c_int64, _P, _P, c_int64, _P]),
    "ngnn_sage_agg_bwd": (c_int32, [_P, _P, _P, c_int64, c_int64, c_int64, _P, c_int64, c_int64, _P, c_int64,
                                    c_float, _P, c_int64, _P]),
    "ngnn_sage_gemm_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "ngnn_sage_gemm_fwd": (c_int32, [_P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int64, c_int64, c_int32,
                                     c_float, c_uint64, c_uint64, _P, c_int64, _P, _P, c_size_t, _P]),
    "ngnn_set_gemm_path": (c_int32, [c_int32]),
    "ngnn_debug_set_trace": (c_int32, [_P]),
    "ngnn_sage_dgrad_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "ngnn_sage_dgrad": (c_int32, [_P, c_int64, _P, _P, _P, c_int64, c_int64, c_int64, _P, c_int64, _P, c_int64, _P,
                                  c_size_t, _P]),
    "ngnn_sage_wgrad_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "ngnn_sage_wgrad": (c_int32, [_P, c_int64, _P, c_int64, _P, c_int64, c_int64, c_int64, c_int64, _P, _P, _P,
                                  c_int32, _P, c_size_t, _P]),
    "ngnn_act_bwd": (c_int32, [_P, c_int64, _P, c_int64, c_int64, c_int64, c_float, _P, c_int64, _P]),
    "ngnn_ce_fwd_bwd": (c_int32, [_P, c_int64, _P, _P, c_int64, c_int64, c_float, _P, _P, c_int64, _P, _P]),
    "ngnn_ce_fwd_bwd_gather": (c_int32, [_P, c_int64, _P, _P, _P, c_int64, c_int64, c_float, _P, _P, c_int64, _P, _P]),
    "ngnn_sage_num_params": (c_int64, [_P]),
    "ngnn_sage_step_workspace_bytes": (c_size_t, [_P, c_int32, _P, _P]),
    "ngnn_sage_step": (c_int32, [_P, _P, _P, _P, _P, _P, _P, c_int64, _P, _P, c_uint64, c_uint64, _P, _P, c_int64, _P,
                                 c_size_t, _P]),
    "ngnn_sample_block_ex": (c_int32, [_P, _P, c_int64, _P, c_int32, _P, c_int32, c_int32, c_uint64, c_uint32, c_uint32, _P,
                                       _P, _P, _P, _P, _P, _P, _P, c_int32, _P, _P, _P, c_size_t, _P]),
    "ngnn_step_ctl_set": (c_int32, [_P, c_uint32, c_uint32, c_uint64, c_float, _P]),
    "ngnn_sage_prep_weights": (c_int32, [_P, _P, c_int32, _P, _P, _P, c_size_t, _P]),
    "ngnn_sage_agg1": (c_int32, [_P, _P, _P, _P, _P, c_int64, c_int32, _P, c_size_t, _P]),
    "ngnn_block_table_index": (c_int32, [_P, _P, _P, _P, c_int32, c_int64, c_int64, _P, _P, _P]),
    "ngnn_sage_forward": (c_int32, [_P, _P, _P, _P, _P, _P, c_int64, c_uint64, c_uint64, _P, c_int64, _P, c_size_t, _P]),
    "ngnn_sage_backward": (c_int32, [_P, _P, _P, _P, _P, _P, _P, c_int64, _P, c_int64, _P, c_size_t, _P]),
    "ngnn_ct_loss": (c_int32, [_P, c_int64, _P, c_int64, _P, _P, _P, _P, c_int64, c_int64, c_int64, _P, _P, c_int64, _P,
                               c_int64, _P, _P, _P, _P]),
    "ngnn_noise_add_fwd": (c_int32, [_P, c_int64, _P, c_int64, _P, c_int64, c_int64, c_float, c_int32, _P, c_int64, _P]),
    "ngnn_noise_add_bwd": (c_int32, [_P, c_int64, _P, c_int64, _P, c_int64, _P, c_int64, c_int64, c_float, c_int32, _P, c_int64, _P]),
    "ngnn_shuffle_rows": (c_int32, [_P, c_int64, c_int64, c_int64, c_int32, c_uint64, c_uint64, _P, c_int64, _P]),
    "ngnn_set_step_overlap": (c_int32, [c_int32]),
    "ngnn_probe_enable": (c_int32, [c_int32]),
    "ngnn_probe_read": (c_int32, [_P, c_int32, _P]),
    "ngnn_probe_read_device_clock": (c_int32, [_P, c_int32, _P]),
    "ngnn_probe_read_agg_t": (c_int32, [_P, c_int32, _P]),
    "ngnn_adam_step": (c_int32, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, c_float, _P,
                                 c_int32, _P]),
    "ngnn_sample_capacity": (c_int32, [c_int32, _P, c_int32, c_int64, _P, _P]),
    "ngnn_sample_workspace_bytes": (c_size_t, [c_int64, c_int32, _P, c_int32]),
    "ngnn_sample_workspace_init": (c_int32, [_P, c_size_t, c_int64, _P]),
    "ngnn_sample_block": (c_int32, [_P, _P, c_int64, _P, c_int32, _P, c_int32, c_int32, c_uint64, c_uint32, c_uint32,
                                    _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
}



class SageModel(ctypes.Structure):      # ngnn_sage_model_t
    _fields_ = [("num_layers", c_int32), ("in_dim", c_int32), ("hidden_dim", c_int32), ("out_dim", c_int32),
                ("dropout", c_float), ("training", c_int32)]


class BlockDesc(ctypes.Structure):      # ngnn_block_t
    _fields_ = [("rowptr", c_void_p), ("col", c_void_p), ("col_global", c_void_p), ("n_id", c_void_p),
                ("num_hops", c_int32), ("hop_nodes", ctypes.POINTER(c_int32)), ("hop_edges", ctypes.POINTER(c_int32)),
                ("colptr_t", c_void_p * 8), ("row_t", c_void_p * 8),
                ("col_table", c_void_p), ("n_table", c_void_p), ("hot_rows", c_int64),
                ("counts", c_void_p), ("batch_size", c_int32), ("ctl", c_void_p), ("agg1_buffer", c_int32),
                ("weights_prepared", c_int32)]


_lib = None


class NgnnError(RuntimeError):
    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed with code {code}: {msg}")
        self.code = code


def lib_path() -> Path:
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load (building first if needed) libngnn_b200.so and attach the signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and _build.needs_build():
        _build.build()
    path = lib_path()
    if not path.exists():
        raise RuntimeError(f"{path} is missing and could not be built; the CUDA path has no fallback")
    lib = ctypes.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError => header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    buf = ctypes.create_string_buffer(512)
    load().ngnn_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(fn: str, rc: int) -> None:
    if rc != 0:
        raise NgnnError(fn, rc, last_error())


def call(fn: str, *args):
    """Call an int32-returning entry point and raise NgnnError on a non-zero code."""
    rc = getattr(load(), fn)(*args)
    if rc != 0:
        raise NgnnError(fn, rc, last_error())
