"""ctypes binding of the host simulation tests/cpu_sim/pg_sim.cpp (test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from bart_rs_b200 import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "cpu_sim", "pg_sim.cpp")
_LIB = os.path.join(_HERE, "cpu_sim", "libpg_sim.so")
_CSRC = os.path.join(os.path.dirname(_HERE), "bart_rs_b200", "csrc")
_lib = None


def build():
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in ("pg_serial.h", "pg_rng.h", "pg_types.h")]
    if os.path.exists(_LIB) and all(os.path.getmtime(_LIB) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["g++", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared",
                           "-o", _LIB, _SRC])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        vp = C.c_void_p
        L.pgsim_create.restype = vp
        L.pgsim_create.argtypes = [vp, C.c_uint64, C.c_uint64, vp, C.POINTER(_capi.Settings), C.c_int]
        L.pgsim_destroy.argtypes = [vp]
        L.pgsim_last_error.restype = C.c_char_p
        L.pgsim_last_error.argtypes = [vp]
        L.pgsim_set_likelihood.argtypes = [vp, C.c_int, vp, C.c_int]
        L.pgsim_set_pg_correct.argtypes = [vp, C.c_int]
        L.pgsim_set_rng_philox.argtypes = [vp, C.c_uint64]
        L.pgsim_set_rng_tape.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int]
        L.pgsim_trace_attach.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int]
        L.pgsim_trace_count.restype = C.c_int64
        L.pgsim_trace_count.argtypes = [vp]
        L.pgsim_trace_overflow.argtypes = [vp]
        L.pgsim_step.argtypes = [vp, C.c_int, vp]
        L.pgsim_get_info.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), vp, C.POINTER(C.c_int)]
        L.pgsim_tree_size.argtypes = [vp, C.c_uint64]
        L.pgsim_get_tree.argtypes = [vp, C.c_uint64, vp, vp, vp]
        L.pgsim_get_leaf_indices.argtypes = [vp, C.c_uint64, vp]
        L.pgsim_get_counters.argtypes = [vp, vp]
        L.pgsim_thr32.restype = C.c_float
        L.pgsim_thr32.argtypes = [C.c_double]
        L.pgsim_f2key.restype = C.c_uint32
        L.pgsim_f2key.argtypes = [C.c_float]
        L.pgsim_key2f.restype = C.c_float
        L.pgsim_key2f.argtypes = [C.c_uint32]
        L.pgsim_bern_terms.argtypes = [vp, vp, vp, vp, C.c_int64]
        L.pgsim_log_ge1.argtypes = [vp, vp, C.c_int64]
        L.pgsim_warp_range.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class SimSampler:
    """Same method names as bart_rs_b200.PySampler's test-facing surface."""

    def __init__(self, x, y, *, n_trees, n_particles, max_depth, alpha, beta, sigma, split_prior=(), batch_tune=0.1,
                 batch_post=0.1, seed=0, grid=3, pg_correct=False, onehot=None):
        L = lib()
        self._L = L
        xf = np.asfortranarray(np.asarray(x, dtype=np.float64))
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
        self.n, self.p = xf.shape
        self.n_trees, self.n_particles = n_trees, n_particles
        prior = np.ascontiguousarray(np.asarray(split_prior, dtype=np.float64))
        names = [b"OneHotSplit" if (onehot is not None and onehot[j]) else b"ContinuousSplit" for j in range(self.p)]
        rules = (C.c_char_p * self.p)(*names)
        st = _capi.Settings(n_trees, n_particles, max_depth, alpha, beta, sigma,
                            prior.ctypes.data_as(C.POINTER(C.c_double)), prior.size, rules, self.p, b"", b"",
                            batch_tune, batch_post, seed)
        self._h = L.pgsim_create(_p(xf), self.n, self.p, _p(y), C.byref(st), grid)
        self._keep = [rules]
        if pg_correct:
            L.pgsim_set_pg_correct(self._h, 1)

    def set_likelihood_code(self, family, params=()):
        a = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
        self._L.pgsim_set_likelihood(self._h, family, _p(a), a.size)

    def set_rng_philox(self, seed):
        self._L.pgsim_set_rng_philox(self._h, seed)

    def set_rng_tape(self, slot, round_, upd):
        slot = np.ascontiguousarray(slot)
        self._L.pgsim_set_rng_tape(self._h, _p(slot), _p(np.ascontiguousarray(round_)), _p(np.ascontiguousarray(upd)),
                                   slot.shape[0], slot.shape[1])

    def trace_enable(self, u_cap, r_cap):
        S, P = self.n_particles - 1, self.n_particles
        self._tr = (np.full((u_cap, r_cap, S, 12), np.nan), np.zeros((u_cap, r_cap, 2 + 2 * S)), np.zeros((u_cap, 4 + 2 * P)))
        self._L.pgsim_trace_attach(self._h, _p(self._tr[0]), _p(self._tr[1]), _p(self._tr[2]), u_cap, r_cap)

    def trace_read(self):
        return (*self._tr, self._L.pgsim_trace_count(self._h), bool(self._L.pgsim_trace_overflow(self._h)))

    def step(self, tune=None):
        out = np.empty(self.n)
        rc = self._L.pgsim_step(self._h, -1 if tune is None else int(bool(tune)), _p(out))
        if rc != 0:
            raise RuntimeError(self._L.pgsim_last_error(self._h).decode())
        return out

    def info(self):
        ll, acc, nd = C.c_double(), C.c_int64(), C.c_int(self.n_trees)
        depths = np.zeros(self.n_trees, dtype=np.uint8)
        self._L.pgsim_get_info(self._h, C.byref(ll), C.byref(acc), _p(depths), C.byref(nd))
        return ll.value, acc.value, depths[: nd.value].copy()

    def tree(self, t):
        size = self._L.pgsim_tree_size(self._h, t)
        sv, sp, lv = np.zeros(size, dtype=np.int32), np.zeros(size), np.zeros(size)
        self._L.pgsim_get_tree(self._h, t, _p(sv), _p(sp), _p(lv))
        return sv, sp, lv

    def leaf_indices(self, t):
        out = np.zeros(self.n, dtype=np.uint32)
        self._L.pgsim_get_leaf_indices(self._h, t, _p(out))
        return out

    def counters(self):
        out = np.zeros(4)
        self._L.pgsim_get_counters(self._h, _p(out))
        return dict(zip(["bytes_model", "bytes_contract", "grid_passes", "tree_updates"], out.tolist()))

    def close(self):
        if self._h:
            self._L.pgsim_destroy(self._h)
            self._h = None
