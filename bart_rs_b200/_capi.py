"""ctypes binding of the C ABI in include/pgbart.h (libpgbart_b200.so).

The library is hand-written CUDA for sm_100a; there is no CPU fallback.  Loading fails loudly when
the shared object has not been built (run `python -c "import __graft_entry__ as g; g.build()"`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGBART_LIB_PATH") or os.path.join(_HERE, "libpgbart_b200.so")   # override: A/B runs of two builds

GAUSSIAN, BERNOULLI_LOGIT, GAUSSIAN_KERNEL = 0, 1, 2
X_F64, X_F32 = 0, 1
ALLOW_FP32_ROUNDING = 0x100
ORDER_C, ORDER_F = 0, 1
TS_FIELDS = 12

EXPORTS = [
    "pgbart_create", "pgbart_destroy", "pgbart_last_error", "pgbart_set_likelihood", "pgbart_step",
    "pgbart_step_device", "pgbart_sync", "pgbart_predictions_device", "pgbart_get_predictions", "pgbart_get_info",
    "pgbart_tree_size", "pgbart_get_tree", "pgbart_get_leaf_indices", "pgbart_get_state", "pgbart_set_rng_philox",
    "pgbart_set_rng_tape", "pgbart_trace_enable", "pgbart_trace_read", "pgbart_set_option", "pgbart_get_counters",
    "pgbart_timer_start", "pgbart_timer_stop", "pgbart_get_profile", "pgbart_debug_block_cycles", "pgbart_debug_state", "pgbart_get_early_stats_counters",
    "pgbart_shard_config", "pgbart_shard_mailbox", "pgbart_shard_connect", "pgbart_predict", "pgbart_variable_inclusion",
    "pgbart_get_spec_counters", "pgbart_host_alloc", "pgbart_host_free",
]


class Settings(C.Structure):
    """struct pgbart_settings (include/pgbart.h) == PyBartSettings (reference src/lib.rs:35-104)."""
    _fields_ = [
        ("n_trees", C.c_uint64), ("n_particles", C.c_uint64), ("max_depth", C.c_uint32),
        ("alpha", C.c_double), ("beta", C.c_double), ("sigma", C.c_double),
        ("split_prior", C.POINTER(C.c_double)), ("n_split_prior", C.c_uint64),
        ("split_rules", C.POINTER(C.c_char_p)), ("n_split_rules", C.c_uint64),
        ("response_rule", C.c_char_p), ("resampling_rule", C.c_char_p),
        ("batch_tune", C.c_double), ("batch_post", C.c_double), ("seed", C.c_uint64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. "
            "This package has no CPU fallback; build it with __graft_entry__.build().")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
    L.pgbart_create.argtypes = [vp, i32, i32, u64, u64, vp, C.POINTER(Settings), i32, C.POINTER(vp)]
    L.pgbart_destroy.argtypes = [vp]
    L.pgbart_last_error.argtypes = [vp]
    L.pgbart_last_error.restype = C.c_char_p
    L.pgbart_set_likelihood.argtypes = [vp, i32, vp, i32]
    L.pgbart_step.argtypes = [vp, i32, vp]
    L.pgbart_step_device.argtypes = [vp, i32]
    L.pgbart_sync.argtypes = [vp]
    L.pgbart_predictions_device.argtypes = [vp]
    L.pgbart_predictions_device.restype = vp
    L.pgbart_get_predictions.argtypes = [vp, vp]
    L.pgbart_get_info.argtypes = [vp, C.POINTER(dbl), C.POINTER(i64), vp, C.POINTER(i32)]
    L.pgbart_tree_size.argtypes = [vp, u64, C.POINTER(i32)]
    L.pgbart_get_tree.argtypes = [vp, u64, vp, vp, vp]
    L.pgbart_get_leaf_indices.argtypes = [vp, u64, vp]
    L.pgbart_get_state.argtypes = [vp, C.POINTER(u64), C.POINTER(i32), C.POINTER(u64)]
    L.pgbart_set_rng_philox.argtypes = [vp, u64]
    L.pgbart_set_rng_tape.argtypes = [vp, vp, vp, vp, i64, i32]
    L.pgbart_trace_enable.argtypes = [vp, i64, i32]
    L.pgbart_trace_read.argtypes = [vp, vp, vp, vp, C.POINTER(i64), C.POINTER(i32)]
    L.pgbart_set_option.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.pgbart_get_counters.argtypes = [vp, vp]
    L.pgbart_timer_start.argtypes = [vp]
    L.pgbart_get_profile.argtypes = [vp, vp]
    L.pgbart_timer_stop.argtypes = [vp, C.POINTER(dbl)]
    L.pgbart_debug_block_cycles.argtypes = [vp, vp, i32]
    L.pgbart_debug_state.argtypes = [vp, vp]
    L.pgbart_get_early_stats_counters.argtypes = [vp, vp]
    L.pgbart_shard_config.argtypes = [vp, i32, i32, u64, dbl, vp, vp]
    L.pgbart_shard_mailbox.argtypes = [vp, vp]
    L.pgbart_shard_connect.argtypes = [vp, vp]
    L.pgbart_predict.argtypes = [vp, vp, i32, i32, u64, vp]
    L.pgbart_variable_inclusion.argtypes = [vp, vp]
    L.pgbart_get_spec_counters.argtypes = [vp, vp]
    L.pgbart_host_alloc.argtypes = [u64, C.POINTER(vp)]
    L.pgbart_host_free.argtypes = [vp]
    _lib = L
    return L
