"""ctypes binding of the C ABI in include/vtc_b200.h (vtc_b200/lib/libvtc_b200.so).

There is no CPU or eager fallback: if the library is missing or fails to load, every op raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvtc_b200.so")

# enums of include/vtc_b200.h
F32, BF16 = 0, 1
METRIC_DOT, METRIC_L2 = 0, 1
PREC_EXACT, PREC_BF16, PREC_BRUTE = 0, 1, 2
OP_SIM_RANK, OP_SIM_TOPK, OP_INFONCE_FWD, OP_SIM_MATRIX, OP_INFONCE_BWD, OP_GT_SCORES, OP_LINEAR = range(7)
CAM_READOUT_AVG, CAM_READOUT_RESIDUAL_ONLY, CAM_READOUT_UNIFORM = 0, 1, 2
RESACT_NONE, RESACT_NORMALIZE_EPS, RESACT_SQUASH, RESACT_TANH, RESACT_AFFINE = range(5)
ABI_VERSION = 3

_P = c_void_p

# name -> (restype, argtypes); kept in the order of the header so that tests can diff them
SIGNATURES = {
    "vtc_abi_version": (c_int, []),
    "vtc_strerror": (c_char_p, [c_int]),
    "vtc_workspace_bytes": (c_size_t, [c_int, c_int64, c_int64, c_int, c_int]),
    "vtc_row_norms": (c_int, [_P, c_int64, c_int, c_int64, c_int, _P, _P, _P]),
    "vtc_normalize": (c_int, [_P, c_int64, c_int, c_int64, c_int, _P, c_int64, _P]),
    "vtc_sim_matrix": (c_int, [_P, _P, c_int64, c_int64, c_int, c_int, c_int, _P, _P, c_int64, _P,
                               c_size_t, _P]),
    "vtc_sim_rank": (c_int, [_P, _P, c_int64, c_int64, c_int, c_int, _P, c_int64, c_int64, c_int,
                             c_int, _P, _P, c_int, _P, _P, c_size_t, _P]),
    "vtc_rank_eval": (c_int, [_P, _P, c_int64, c_int64, c_int, c_int, _P, c_int, c_int, POINTER(c_int),
                              c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "vtc_rank_prepare": (c_int, [_P, c_int64, c_int, c_int, c_int, _P, _P, _P]),
    "vtc_sim_rank_prepared": (c_int, [_P, _P, c_int64, c_int64, c_int, c_int, _P, c_int64, c_int64,
                                      c_int, c_int, _P, _P, _P, _P, _P, _P, c_int, _P, _P, c_size_t,
                                      _P]),
    "vtc_gt_scores": (c_int, [_P, _P, c_int64, c_int64, c_int, c_int, _P, c_int64, c_int64, c_int,
                              c_int, _P, _P, c_size_t, _P]),
    "vtc_rank_finalize": (c_int, [_P, _P, c_int64, c_int64, POINTER(c_int), c_int, _P, _P, _P,
                                  c_size_t, _P]),
    "vtc_sim_topk": (c_int, [_P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, c_int64,
                             _P, _P, _P, c_size_t, _P]),
    "vtc_topk_merge": (c_int, [_P, _P, c_int, c_int64, c_int, _P, _P, _P]),
    "vtc_infonce_fwd": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P,
                                c_size_t, _P]),
    "vtc_infonce_bwd": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P,
                                c_size_t, _P]),
    "vtc_infonce_dense_fwd": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "vtc_infonce_dense_bwd": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, c_int64, _P]),
    "vtc_cam_stack_normalize": (c_int, [_P, _P, c_int, c_int64, c_int, _P, _P]),
    "vtc_layernorm": (c_int, [_P, _P, _P, c_int64, c_int, c_float, _P, _P]),
    "vtc_cam_attn_core": (c_int, [_P, c_int, c_int64, c_int, c_int, _P, _P]),
    "vtc_bias_act": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, _P]),
    "vtc_cam_readout": (c_int, [_P, _P, _P, _P, c_int, c_int64, c_int, c_int, c_int, c_float, _P, _P, _P,
                                _P]),
    "vtc_linear": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P, _P, c_size_t,
                           _P]),
    "vtc_cam_backward_workspace_bytes": (c_size_t, [c_int, c_int64, c_int, c_int]),
    "vtc_cam_backward": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int64, c_int, c_int, c_int, _P, c_int,
                                 _P, c_int, c_float, _P, _P, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "vtc_launch_count": (c_uint64, []),
    "vtc_kernel_timer_enable": (c_int, [c_int]),
    "vtc_kernel_timer_read": (c_int, [POINTER(c_double), POINTER(c_int)]),
    "vtc_debug_prof_read": (c_int, [POINTER(c_uint64), c_int]),
    "vtc_trace_begin": (c_int, [_P]),
    "vtc_trace_end": (c_int, [ctypes.c_char_p, c_size_t]),
}


class CamLayerBwd(ctypes.Structure):
    """vtc_cam_layer_bwd of include/vtc_b200.h (device pointers)."""
    _fields_ = [(n, c_void_p) for n in (
        "X", "H1", "QKV", "A", "X2", "H2", "U", "Fa", "ln1_g", "ln2_g", "qkv_t", "out_t", "fc_t",
        "proj_t", "dWqkv", "dbqkv", "dWo", "dbo", "dg1", "db1", "dWfc", "dbfc", "dWpr", "dbpr",
        "dg2", "db2")]


class CamLayer(ctypes.Structure):
    """vtc_cam_layer of include/vtc_b200.h (device pointers)."""
    _fields_ = [(n, c_void_p) for n in ("ln1_g", "ln1_b", "ln2_g", "ln2_b", "qkv", "out", "fc", "proj")]


SIGNATURES.update({
    "vtc_linear_prepared_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vtc_linear_prepare": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P]),
    "vtc_cam_workspace_bytes": (c_size_t, [c_int, c_int64, c_int, c_int]),
    "vtc_cam_forward": (c_int, [_P, _P, c_int, c_int64, c_int, c_int, c_int, POINTER(CamLayer), c_int,
                                _P, _P, c_int, c_float, _P, _P, c_int, _P, _P, c_size_t, _P]),
})


SIGNATURES.update({
    "vtc_normalize_bwd": (c_int, [_P, _P, c_int64, c_int, _P, _P]),
    "vtc_transpose": (c_int, [_P, c_int64, c_int64, _P, _P]),
    "vtc_gelu_bwd": (c_int, [_P, _P, c_int64, _P, _P]),
    "vtc_colsum": (c_int, [_P, c_int64, c_int64, _P, _P]),
    "vtc_layernorm_bwd": (c_int, [_P, _P, _P, c_int64, c_int, c_float, _P, _P, _P, _P, _P]),
    "vtc_cam_attn_core_bwd": (c_int, [_P, _P, c_int, c_int64, c_int, c_int, _P, _P]),
    "vtc_cam_stack_normalize_bwd": (c_int, [_P, _P, _P, c_int, c_int64, c_int, _P, _P, _P]),
    "vtc_cam_readout_bwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int64, c_int, c_int, c_int, c_float,
                                    _P, _P, _P, _P, _P, _P]),
})


class VtcError(RuntimeError):
    pass


_lib = None


def load(path: str = LIB_PATH) -> ctypes.CDLL:
    """Load the library and bind every symbol the header declares.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise VtcError(
            f"{path} not found: build it with `python -m vtc_b200.build` (needs nvcc, sm_100a). "
            "vtc_b200 has no CPU or eager fallback.")
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = the .so is stale
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.vtc_abi_version()
    if got != ABI_VERSION:
        raise VtcError(f"libvtc_b200.so ABI {got} != binding ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().vtc_strerror(rc)
        raise VtcError(f"{what} failed: {msg.decode() if msg else rc} ({rc})")


def launch_count() -> int:
    return int(load().vtc_launch_count())


def kernel_timer_enable(on: bool) -> None:
    check(load().vtc_kernel_timer_enable(1 if on else 0), "vtc_kernel_timer_enable")


def kernel_timer_read():
    """(total milliseconds, launches) of the tensor-core launches since the last read."""
    ms = c_double(0.0)
    n = c_int(0)
    check(load().vtc_kernel_timer_read(ctypes.byref(ms), ctypes.byref(n)), "vtc_kernel_timer_read")
    return ms.value, n.value


def debug_prof_read(n_ctas: int = 148):
    """VTC_DBG_PROF=1 only: per-CTA wait counters of the last tensor-core launch, as a list of
    8-tuples (see include/vtc_b200.h::vtc_debug_prof_read); synchronises the device."""
    buf = (c_uint64 * (8 * n_ctas))()
    check(load().vtc_debug_prof_read(buf, 8 * n_ctas), "vtc_debug_prof_read")
    return [tuple(int(buf[8 * i + j]) for j in range(8)) for i in range(n_ctas)]


def trace_begin(stream: int = 0) -> None:
    """Profiling aid: start recording one CUDA event per library launch on `stream` (a raw handle)."""
    check(load().vtc_trace_begin(stream), "vtc_trace_begin")


def trace_end(cap: int = 1 << 18):
    """Stop the launch trace; returns [(source "file:line", microseconds)] in launch order."""
    buf = ctypes.create_string_buffer(cap)
    n = load().vtc_trace_end(buf, cap)
    if n < 0:
        check(n, "vtc_trace_end")
    out = []
    for ln in buf.value.decode().splitlines():
        where, us = ln.rsplit(" ", 1)
        out.append((where, float(us)))
    return out
