"""ctypes binding of libvbn_cuda.so (include/vbn_cuda.h).  There is no CPU fallback: if the
library is missing or no CUDA device is present the product path raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libvbn_cuda.so")

ABI_VERSION = 6

# op kinds / flags (keep in sync with include/vbn_cuda.h)
OP_NONE, OP_LG, OP_GNN, OP_MDN, OP_SNN, OP_KDE, OP_TAB, OP_RFF = 0, 1, 2, 3, 4, 5, 6, 7
OP_TAKEW, OP_SELECT, OP_JUMP = 8, 9, 10
SRC_SAMPLE, SRC_FIXED_Q, SRC_FIXED_ROW, SRC_SLOT = 0, 1, 2, 3
F_ADD_LOGW, F_OUT_LOGP, F_SHARED, F_FAST32, F_LGFAST, F_LGPLAIN = 0x4, 0x8, 0x10, 0x20, 0x40, 0x80
F_PAR4, F_MDNPLAIN, F_OUT_PARAMS, F_MDNROOT, F_TABPLAIN, F_MDNFAST = 0x100, 0x200, 0x400, 0x800, 0x1000, 0x2000
ACT = {"relu": 0, "tanh": 1, "gelu": 2, "elu": 3}
WITHIN_BIN = {"uniform": 0, "triangular": 1, "gaussian": 2}
MAX_LAYERS = 8
MAX_GENERIC_WIDTH = 128
TC_WBUF_BYTES = 30720  # one slot of the tensor-core kernel's weight ring (csrc/vbn_tc_layout.h kWbufBytes)

OP_DTYPE = np.dtype(
    [
        ("kind", "<i4"), ("flags", "<i4"), ("dim", "<i4"), ("n_par", "<i4"),
        ("out_slot", "<i4"), ("par_off", "<i4"), ("param_off", "<i4"), ("fixed_col", "<i4"),
        ("store_idx", "<i4"), ("noise_idx", "<i4"), ("n_off", "<i4"), ("u_off", "<i4"),
        ("n_layers", "<i4"), ("act", "<i4"), ("n_out", "<i4"), ("k", "<i4"),
        ("layer_dim", "<i4", (MAX_LAYERS,)), ("aux", "<i4", (4,)), ("tc", "<i4", (4,)),
    ]
)
assert OP_DTYPE.itemsize == 128


class ProgramDesc(C.Structure):
    _fields_ = [
        ("ops_dev", C.c_void_p), ("n_ops", C.c_int32),
        ("par_slots_dev", C.c_void_p), ("n_par_slots", C.c_int32),
        ("params_dev", C.c_void_p), ("n_params", C.c_int64),
        ("n_slots", C.c_int32), ("n_scratch", C.c_int32),
        ("heavy", C.c_int32), ("tc", C.c_int32),
        ("tc_list_dev", C.c_void_p), ("n_tc", C.c_int32), ("tc_image_bytes", C.c_int32),
        ("has_tables", C.c_int32), ("rows_per_thread", C.c_int32),
    ]


class RunDesc(C.Structure):
    _fields_ = [
        ("n_queries", C.c_int64), ("n_samples", C.c_int64),
        ("query_offset", C.c_int64), ("sample_offset", C.c_int64),
        ("seed", C.c_uint64), ("call_offset", C.c_uint64),
        ("fixed_dev", C.c_void_p), ("inputs_dev", C.c_void_p),
        ("stores_dev", C.c_void_p), ("noise_dev", C.c_void_p),
        ("logw_dev", C.c_void_p), ("logp_dev", C.c_void_p),
        ("logp_as_pdf", C.c_int32), ("logw_accumulate", C.c_int32),
        ("error_flag_dev", C.c_void_p),
        ("seg_dev", C.c_void_p), ("seg_per_query", C.c_int32), ("seg_slot", C.c_int32), ("seg_classes", C.c_int32),
    ]


EXPORTS = {
    "vbn_cuda_abi_version": (C.c_int32, []),
    "vbn_cuda_last_error": (C.c_char_p, []),
    "vbn_cuda_device_count": (C.c_int32, [C.POINTER(C.c_int32)]),
    "vbn_plan_create": (C.c_int32, [C.POINTER(ProgramDesc), C.POINTER(C.c_void_p)]),
    "vbn_plan_destroy": (C.c_int32, [C.c_void_p]),
    "vbn_run_forward": (C.c_int32, [C.c_void_p, C.POINTER(RunDesc), C.c_void_p]),
    "vbn_run_forward_launches": (C.c_int32, [C.c_void_p]),
    "vbn_lse_partials": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "vbn_lse_merge": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "vbn_weights_normalize": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                          C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vbn_segment_merge": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "vbn_ess_below": (C.c_int32, [C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]),
    "vbn_row_cdf": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "vbn_resample_indices": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, C.c_int64,
                                         C.c_int64, C.c_void_p, C.c_void_p]),
    "vbn_gather_rows": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p]),
    "vbn_weighted_histogram": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                           C.c_void_p, C.c_void_p]),
    "vbn_weighted_sum": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "vbn_gaussian_mixture_grid": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_float,
                                              C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vbn_gaussian_grid": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "vbn_posterior_stats": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                        C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vbn_kde_log_prob": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float,
                                     C.c_float, C.c_void_p, C.c_void_p]),
    "vbn_fma_peak": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "vbn_tf32_peak": (C.c_int32, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "vbn_kde_tc_workspace_bytes": (C.c_int32, [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]),
    "vbn_kde_log_prob_tc": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "vbn_philox_fill": (C.c_int32, [C.c_void_p, C.c_int64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "vbn_stream_draws": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                     C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
}

_lib = None


class VbnCudaError(RuntimeError):
    pass


def load():
    """Loads the shared library (no GPU needed to load it) and binds every export."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VbnCudaError(
            f"{LIB_PATH} not found: build it with `python -m vectorizedbayesiannetwork_b200.build` "
            "(there is no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.vbn_cuda_abi_version() != ABI_VERSION:
        raise VbnCudaError("libvbn_cuda.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        msg = load().vbn_cuda_last_error()
        raise VbnCudaError(f"libvbn_cuda error {code}: {msg.decode() if msg else ''}")


def launch_count() -> int:
    return _LAUNCHES[0]


_LAUNCHES = [0]


def count_launch(n: int = 1) -> None:
    _LAUNCHES[0] += n
