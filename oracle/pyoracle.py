"""ctypes binding of the CPU oracle (oracle/pgbart_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(bart_rs_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpgbart_oracle.so")

FAM_GAUSSIAN, FAM_BERNOULLI_LOGIT, FAM_GAUSSIAN_KERNEL = 0, 1, 2
RNG_SEQ, RNG_PHILOX, RNG_REPLAY, RNG_FLAT = 0, 1, 2, 3
TS_FIELDS = 12


class _Settings(C.Structure):
    _fields_ = [
        ("n_trees", C.c_uint64), ("n_particles", C.c_uint64), ("max_depth", C.c_uint32),
        ("alpha", C.c_double), ("beta", C.c_double), ("sigma", C.c_double),
        ("split_prior", C.POINTER(C.c_double)), ("n_split_prior", C.c_uint64),
        ("batch_tune", C.c_double), ("batch_post", C.c_double), ("seed", C.c_uint64),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pgbart_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(_LIB_PATH)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_LIB_PATH)
        L = _lib
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.POINTER(_Settings)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_likelihood.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_set_modes.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_set_rng.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.orc_attach_tape.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.orc_attach_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.orc_attach_flat.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_record_flat.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_flat_count.restype = C.c_uint64
        L.orc_flat_count.argtypes = [C.c_void_p]
        L.orc_trace_count.restype = C.c_int64
        L.orc_trace_count.argtypes = [C.c_void_p]
        L.orc_overflowed.argtypes = [C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_get_info.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_void_p,
                                   C.POINTER(C.c_int)]
        L.orc_tree_size.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_get_tree.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_leaf_indices.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_next_tree_idx.restype = C.c_uint64
        L.orc_next_tree_idx.argtypes = [C.c_void_p]
        L.orc_upd_counter.restype = C.c_uint64
        L.orc_upd_counter.argtypes = [C.c_void_p]
        L.orc_kat_max_nodes_for_depth.restype = C.c_uint64
        L.orc_kat_max_nodes_for_depth.argtypes = [C.c_uint32]
        L.orc_kat_get_depth.restype = C.c_uint64
        L.orc_kat_get_depth.argtypes = [C.c_uint64]
        L.orc_kat_tree.argtypes = [C.c_double, C.c_uint64, C.c_uint32, C.c_int] + [C.c_void_p] * 9 + [C.c_int]
        L.orc_kat_apply_mutation.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64,
                                             C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int),
                                             C.POINTER(C.c_int)]
        L.orc_kat_normalize.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_kat_systematic.argtypes = [C.c_double, C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_kat_sample_feature.restype = C.c_uint64
        L.orc_kat_sample_feature.argtypes = [C.c_double, C.c_void_p, C.c_uint64]
        L.orc_kat_weighted_index.restype = C.c_uint64
        L.orc_kat_weighted_index.argtypes = [C.c_double, C.c_void_p, C.c_uint64]
        L.orc_kat_xoshiro.argtypes = [C.c_uint64] * 4 + [C.c_int, C.c_void_p]
        L.orc_kat_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_kat_unrolled_mean.restype = C.c_double
        L.orc_kat_unrolled_mean.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_kat_loglik.restype = C.c_double
        L.orc_kat_loglik.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_uint64]
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class TraceBuffers:
    """Trace / tape buffers with the layout documented in include/pgbart.h."""

    def __init__(self, u_cap: int, r_cap: int, n_particles: int):
        S, P = n_particles - 1, n_particles
        self.u_cap, self.r_cap, self.S, self.P = u_cap, r_cap, S, P
        self.slot = np.full((u_cap, r_cap, S, TS_FIELDS), np.nan)
        self.round = np.zeros((u_cap, r_cap, 2 + 2 * S))
        self.upd = np.zeros((u_cap, 4 + 2 * P))


class TapeBuffers:
    def __init__(self, n_upd: int, r_cap: int, n_particles: int):
        S = n_particles - 1
        self.n_upd, self.r_cap, self.S = n_upd, r_cap, S
        self.slot = np.zeros((n_upd, r_cap, S, 5))
        self.round = np.zeros((n_upd, r_cap))
        self.upd = np.zeros((n_upd,))


class OracleSampler:
    """CPU oracle with the PySampler-shaped surface (init/step)."""

    def __init__(self, x, y, *, n_trees, n_particles, max_depth, alpha, beta, sigma, split_prior=(),
                 batch_tune=0.1, batch_post=0.1, seed=0, pg_correct=False, onehot=None):
        L = lib()
        x = np.asarray(x, dtype=np.float64)
        self.n, self.p = x.shape
        self._x = np.asfortranarray(x)
        self._y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
        self._prior = np.ascontiguousarray(np.asarray(split_prior, dtype=np.float64))
        st = _Settings(n_trees, n_particles, max_depth, alpha, beta, sigma,
                       self._prior.ctypes.data_as(C.POINTER(C.c_double)), self._prior.size,
                       batch_tune, batch_post, seed)
        self.n_trees, self.n_particles = n_trees, n_particles
        self._h = L.orc_create(_ptr(self._x), self.n, self.p, _ptr(self._y), C.byref(st))
        self._keep = []
        if pg_correct or onehot is not None:      # SURVEY.md 8(f4) modes (not the reference's behaviour)
            oh = None if onehot is None else np.ascontiguousarray(np.asarray(onehot, dtype=np.uint8))
            assert oh is None or oh.size == self.p
            L.orc_set_modes(self._h, int(bool(pg_correct)), None if oh is None else _ptr(oh))

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_likelihood(self, family, params=()):
        a = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
        lib().orc_set_likelihood(self._h, family, _ptr(a), a.size)

    def set_rng(self, mode, seed=0):
        lib().orc_set_rng(self._h, mode, seed)

    def attach_tape(self, tape: TapeBuffers):
        self._keep.append(tape)
        lib().orc_attach_tape(self._h, _ptr(tape.slot), _ptr(tape.round), _ptr(tape.upd), tape.n_upd, tape.r_cap)

    def attach_trace(self, tr: TraceBuffers):
        self._keep.append(tr)
        lib().orc_attach_trace(self._h, _ptr(tr.slot), _ptr(tr.round), _ptr(tr.upd), tr.u_cap, tr.r_cap)

    def attach_flat(self, log: np.ndarray):
        """RNG_FLAT: consume a sequential log (raw u64 words in float64 bit patterns, normals by value)."""
        log = np.ascontiguousarray(log, dtype=np.float64)
        self._keep.append(log)
        lib().orc_attach_flat(self._h, _ptr(log), log.size)

    def record_flat(self, out: np.ndarray):
        """RNG_SEQ: also write the sequential log RNG_FLAT consumes."""
        self._keep.append(out)
        lib().orc_record_flat(self._h, _ptr(out), out.size)

    def flat_count(self):
        return lib().orc_flat_count(self._h)

    def trace_count(self):
        return lib().orc_trace_count(self._h)

    def overflowed(self):
        return lib().orc_overflowed(self._h)

    def step(self, tune=None):
        out = np.empty(self.n)
        rc = lib().orc_step(self._h, -1 if tune is None else int(bool(tune)), _ptr(out))
        if rc != 0:
            raise RuntimeError(lib().orc_last_error(self._h).decode())
        return out

    def info(self):
        ll, acc, nd = C.c_double(), C.c_int64(), C.c_int(self.n_trees)
        depths = np.zeros(self.n_trees, dtype=np.uint8)
        lib().orc_get_info(self._h, C.byref(ll), C.byref(acc), _ptr(depths), C.byref(nd))
        return ll.value, acc.value, depths[: nd.value].copy()

    def tree(self, t):
        size = lib().orc_tree_size(self._h, t)
        sv = np.zeros(size, dtype=np.int32)
        sp = np.zeros(size)
        lv = np.zeros(size)
        lib().orc_get_tree(self._h, t, _ptr(sv), _ptr(sp), _ptr(lv))
        return sv, sp, lv

    def leaf_indices(self, t):
        out = np.zeros(self.n, dtype=np.uint32)
        lib().orc_get_leaf_indices(self._h, t, _ptr(out))
        return out
