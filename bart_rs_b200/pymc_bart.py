"""Drop-in for the reference's compiled module `pymc_bart.pymc_bart` (src/lib.rs:205-210).

Same two classes, same constructor keywords and method names:

    PyBartSettings(n_trees, n_particles, max_depth, alpha, beta, sigma, split_prior, split_rules,
                   response_rule, resampling_rule, batch_tune=0.1, batch_post=0.1, seed=0)   lib.rs:54-104
    PySampler.init(x, y, model, settings) -> PySampler                                       lib.rs:115-180
    PySampler.step(tune=None) -> float64[n]                                                   lib.rs:182-202

The sweep itself runs in libpgbart_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/pgbart.h).  One difference is forced by keeping the whole sweep on the device: the
reference's `model` argument is the address of a host log-likelihood callback (numba cfunc,
compile_pymc.py:256-307) that is called per particle per round; a host callback cannot stay on the
device, so `model` must be a `DeviceLikelihood` (Gaussian or Bernoulli-logit).  Passing an integer
address raises ValueError: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _capi


@dataclass
class DeviceLikelihood:
    """Likelihood evaluated on the device in place of the reference's LogpFunc pointer (lib.rs:33).

    family: "gaussian" (params: sigma), "bernoulli_logit", or "gaussian_kernel" (the bench's mock,
    benches/particle_gibbs.rs:17-33).
    """
    family: str = "gaussian"
    sigma: float = 1.0

    _CODES = {"gaussian": _capi.GAUSSIAN, "bernoulli_logit": _capi.BERNOULLI_LOGIT,
              "gaussian_kernel": _capi.GAUSSIAN_KERNEL}

    def code(self) -> int:
        try:
            return self._CODES[self.family]
        except KeyError:
            raise ValueError(
                f"likelihood family {self.family!r} is not available on the device path "
                f"(supported: {sorted(self._CODES)}); there is no CPU fallback") from None

    def params(self) -> np.ndarray:
        return np.array([self.sigma], dtype=np.float64) if self.family == "gaussian" else np.zeros(0)


class PyBartSettings:
    """Mirror of `#[pyclass] PyBartSettings` (reference src/lib.rs:35-104): same 13 fields and defaults."""

    def __init__(self, n_trees: int, n_particles: int, max_depth: int, alpha: float, beta: float, sigma: float,
                 split_prior: Sequence[float], split_rules: Sequence[str], response_rule: str,
                 resampling_rule: str, batch_tune: float = 0.1, batch_post: float = 0.1, seed: int = 0):
        self.n_trees = int(n_trees)
        self.n_particles = int(n_particles)
        if not 0 <= int(max_depth) <= 255:
            raise OverflowError("max_depth must fit in u8")  # PyO3 conversion error in the reference
        self.max_depth = int(max_depth)
        self.alpha = float(alpha)
        self.beta = float(beta)
        self.sigma = float(sigma)
        self.split_prior = [float(v) for v in split_prior]
        self.split_rules = [str(r) for r in split_rules]
        self.response_rule = str(response_rule)      # accepted and unused, as in the reference (Q6)
        self.resampling_rule = str(resampling_rule)  # accepted and unused: always systematic (lib.rs:166)
        self.batch_tune = float(batch_tune)
        self.batch_post = float(batch_post)
        self.seed = int(seed)

    def __repr__(self):
        return (f"PyBartSettings(n_trees={self.n_trees}, n_particles={self.n_particles}, max_depth={self.max_depth}, "
                f"alpha={self.alpha}, beta={self.beta}, sigma={self.sigma}, batch_tune={self.batch_tune}, "
                f"batch_post={self.batch_post}, seed={self.seed})")


def not_fp32_exact(a) -> bool:
    """True if some finite value of `a` changes when stored as fp32 (what the device stores; see PySampler.init)."""
    a = np.asarray(a)
    if a.dtype == np.float32:
        return False
    a64 = a.astype(np.float64, copy=False)
    with np.errstate(over="ignore", invalid="ignore"):
        back = a64.astype(np.float32).astype(np.float64)
    return bool(np.any((back != a64) & ~np.isnan(a64)))


class _PinnedPool:
    """Page-locked float64[n] blocks that back the arrays PySampler.step() returns.

    The reference returns a fresh array per call (lib.rs:199).  A fresh pageable array costs a page-faulting allocation
    plus a staged device->host copy (1.3 ms for 8 MB at BASELINE C3); a pinned block takes the copy at link speed.  Each
    returned array owns its block until it is garbage collected, then the block goes back to the pool, so callers still
    see a distinct array per call.  At most `cap` blocks are kept pinned; beyond that step() falls back to np.empty."""

    def __init__(self, L, n: int, cap: int = 8):
        self.L, self.n, self.cap = L, int(n), cap
        self.free, self.lent, self.closed, self.warm = [], 0, False, False

    def take(self):
        import weakref
        if self.closed or self.n == 0:
            return None
        if self.free:
            ptr = self.free.pop()
        elif self.lent >= self.cap:
            return None
        else:
            ptr = self._alloc()
            if ptr is None:
                return None
            if not self.warm:
                # the usual caller keeps the previous step's array alive while it asks for the next one (`pred = s.step()`
                # in a loop): pin the second block now, at the first step, rather than inside the second step (page-locking
                # 8 MB costs ~2 ms, a fifth of a sweep at BASELINE C3)
                self.warm = True
                spare = self._alloc()
                if spare is not None:
                    self.free.append(spare)
        self.lent += 1
        arr = np.ctypeslib.as_array((C.c_double * self.n).from_address(ptr))
        weakref.finalize(arr, self._give, ptr)
        return arr

    def prewarm(self, k: int) -> None:
        """Pin k blocks now (observation-sharded samplers: nothing may be allocated inside a step)."""
        self.warm = True
        while len(self.free) + self.lent < min(k, self.cap):
            ptr = self._alloc()
            if ptr is None:
                break
            self.free.append(ptr)

    def _alloc(self):
        p = C.c_void_p()
        if self.L.pgbart_host_alloc(self.n * 8, C.byref(p)) != 0 or not p.value:
            return None
        return p.value

    def _give(self, ptr):
        self.lent -= 1
        if self.closed:
            self.L.pgbart_host_free(C.c_void_p(ptr))
        else:
            self.free.append(ptr)

    def close(self):
        self.closed = True
        for ptr in self.free:
            self.L.pgbart_host_free(C.c_void_p(ptr))
        self.free = []


class PySampler:
    """Mirror of `#[pyclass(unsendable)] PySampler` (reference src/lib.rs:106-202).

    Use from the thread that created it (one CUDA stream per sampler)."""

    def __init__(self):
        raise TypeError("use PySampler.init(x, y, model, settings)")  # the reference has no #[new]

    @staticmethod
    def init(x, y, model, settings: PyBartSettings, device: int = 0, fp32_rounding: str = "error") -> "PySampler":
        """x, y, model, settings as in the reference (lib.rs:115-180).  The device stores X and y as fp32; float64 data
        that is not exactly representable in fp32 is refused (ValueError) unless fp32_rounding="allow", because the
        reference computes on the f64 values and a silent quantisation can change ties, counts and decisions."""
        if fp32_rounding not in ("error", "allow"):
            raise ValueError("fp32_rounding must be 'error' or 'allow'")
        if isinstance(model, (int, np.integer)):
            raise ValueError(
                "model must be a DeviceLikelihood: the reference's host logp callback (lib.rs:125-126) "
                "cannot run inside the device-resident sweep and this package has no CPU fallback")
        if not isinstance(model, DeviceLikelihood):
            raise TypeError("model must be a DeviceLikelihood")
        L = _capi.lib()
        x = np.asarray(x)
        if x.ndim != 2:
            raise TypeError("x must be a 2-D array")
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
        if y.ndim != 1 or y.shape[0] != x.shape[0]:
            raise ValueError("y must be 1-D with one entry per row of x")
        if x.dtype == np.float32:
            dtype = _capi.X_F32
        else:
            x = x.astype(np.float64, copy=False)
            dtype = _capi.X_F64
        if x.flags.f_contiguous:
            order = _capi.ORDER_F
        else:
            x = np.ascontiguousarray(x)
            order = _capi.ORDER_C
        n, p = x.shape
        prior = np.ascontiguousarray(np.asarray(settings.split_prior, dtype=np.float64))
        rules = (C.c_char_p * max(len(settings.split_rules), 1))(*[r.encode() for r in settings.split_rules])
        st = _capi.Settings(
            settings.n_trees, settings.n_particles, settings.max_depth, settings.alpha, settings.beta, settings.sigma,
            prior.ctypes.data_as(C.POINTER(C.c_double)), prior.size, rules, len(settings.split_rules),
            settings.response_rule.encode(), settings.resampling_rule.encode(),
            settings.batch_tune, settings.batch_post, settings.seed)
        h = C.c_void_p()
        flags = _capi.ALLOW_FP32_ROUNDING if fp32_rounding == "allow" else 0
        rc = L.pgbart_create(x.ctypes.data_as(C.c_void_p), dtype | flags, order, n, p, y.ctypes.data_as(C.c_void_p),
                             C.byref(st), device, C.byref(h))
        if rc != 0:
            msg = L.pgbart_last_error(None).decode()
            raise (ValueError if rc == 1 else RuntimeError)(msg)
        self = object.__new__(PySampler)
        self._h = h
        self._L = L
        self.n, self.p = n, p
        self.n_trees, self.n_particles = settings.n_trees, settings.n_particles
        self._pool = _PinnedPool(L, n)
        self.set_likelihood(model)
        return self

    # -- reference surface ----------------------------------------------------------------------
    def step(self, tune: Optional[bool] = None) -> np.ndarray:
        """One batch of tree updates; returns a fresh float64[n] sum-of-trees (lib.rs:182-202)."""
        out = self._pool.take()
        if out is None:
            out = np.empty(self.n, dtype=np.float64)
        self._check(self._L.pgbart_step(self._h, -1 if tune is None else int(bool(tune)),
                                        out.ctypes.data_as(C.c_void_p)))
        return out

    # -- device-path additions --------------------------------------------------------------------
    def set_likelihood(self, lik: DeviceLikelihood) -> None:
        """Push the current likelihood parameters (the reference's update_shared_arrays, pgbart.py:190)."""
        prm = lik.params()
        self._check(self._L.pgbart_set_likelihood(self._h, lik.code(), prm.ctypes.data_as(C.c_void_p), prm.size))

    def step_device(self, tune: Optional[bool] = None) -> None:
        """Enqueue a step and leave the result in HBM (asynchronous; call sync())."""
        self._check(self._L.pgbart_step_device(self._h, -1 if tune is None else int(bool(tune))))

    def sync(self) -> None:
        self._check(self._L.pgbart_sync(self._h))

    def predictions_device_ptr(self) -> int:
        return int(self._L.pgbart_predictions_device(self._h))

    def predictions(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.float64)
        self._check(self._L.pgbart_get_predictions(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def info(self):
        """BartInfo of the last step (state.rs:16-20): (log_likelihood, acceptance_count, tree_depths)."""
        ll, acc, nd = C.c_double(), C.c_int64(), C.c_int(self.n_trees)
        depths = np.zeros(self.n_trees, dtype=np.uint8)
        self._check(self._L.pgbart_get_info(self._h, C.byref(ll), C.byref(acc), depths.ctypes.data_as(C.c_void_p),
                                            C.byref(nd)))
        return ll.value, acc.value, depths[: nd.value].copy()

    def tree(self, t: int):
        """(split_var int32 with -1 = leaf, split_val, leaf_val) of forest tree t (tree.rs:9-23)."""
        size = C.c_int()
        self._check(self._L.pgbart_tree_size(self._h, t, C.byref(size)))
        sv = np.zeros(size.value, dtype=np.int32)
        sp = np.zeros(size.value)
        lv = np.zeros(size.value)
        self._check(self._L.pgbart_get_tree(self._h, t, sv.ctypes.data_as(C.c_void_p), sp.ctypes.data_as(C.c_void_p),
                                            lv.ctypes.data_as(C.c_void_p)))
        return sv, sp, lv

    def leaf_indices(self, t: int) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.uint32)
        self._check(self._L.pgbart_get_leaf_indices(self._h, t, out.ctypes.data_as(C.c_void_p)))
        return out

    def state(self):
        nxt, tune, nupd = C.c_uint64(), C.c_int(), C.c_uint64()
        self._check(self._L.pgbart_get_state(self._h, C.byref(nxt), C.byref(tune), C.byref(nupd)))
        return nxt.value, bool(tune.value), nupd.value

    def set_rng_philox(self, seed: int) -> None:
        self._check(self._L.pgbart_set_rng_philox(self._h, seed))

    def set_rng_tape(self, slot: np.ndarray, round_: np.ndarray, upd: np.ndarray) -> None:
        slot = np.ascontiguousarray(slot, dtype=np.float64)
        round_ = np.ascontiguousarray(round_, dtype=np.float64)
        upd = np.ascontiguousarray(upd, dtype=np.float64)
        n_upd, r_cap = slot.shape[0], slot.shape[1]
        self._check(self._L.pgbart_set_rng_tape(self._h, slot.ctypes.data_as(C.c_void_p),
                                                round_.ctypes.data_as(C.c_void_p), upd.ctypes.data_as(C.c_void_p),
                                                n_upd, r_cap))

    def trace_enable(self, u_cap: int, r_cap: int) -> None:
        self._trace_shape = (u_cap, r_cap)
        self._check(self._L.pgbart_trace_enable(self._h, u_cap, r_cap))

    def trace_read(self):
        u_cap, r_cap = self._trace_shape
        S, P = self.n_particles - 1, self.n_particles
        slot = np.empty((u_cap, r_cap, S, _capi.TS_FIELDS))
        rnd = np.empty((u_cap, r_cap, 2 + 2 * S))
        upd = np.empty((u_cap, 4 + 2 * P))
        n_upd, ovf = C.c_int64(), C.c_int()
        self._check(self._L.pgbart_trace_read(self._h, slot.ctypes.data_as(C.c_void_p), rnd.ctypes.data_as(C.c_void_p),
                                              upd.ctypes.data_as(C.c_void_p), C.byref(n_upd), C.byref(ovf)))
        return slot, rnd, upd, n_upd.value, bool(ovf.value)

    def set_option(self, key: str, value: str) -> None:
        self._check(self._L.pgbart_set_option(self._h, key.encode(), str(value).encode()))

    def counters(self) -> dict:
        out = np.zeros(8)
        self._check(self._L.pgbart_get_counters(self._h, out.ctypes.data_as(C.c_void_p)))
        keys = ["bytes_model", "bytes_contract", "grid_passes", "tree_updates", "kernel_launches", "last_step_gpu_us",
                "grid_blocks", "smem_bytes"]
        return dict(zip(keys, out.tolist()))

    def spec_counters(self) -> dict:
        """Root passes issued ahead / early-commit predictions / mispredictions / drained passes since creation."""
        out = np.zeros(4, dtype=np.uint64)
        self._check(self._L.pgbart_get_spec_counters(self._h, out.ctypes.data_as(C.c_void_p)))
        return dict(zip(["issued_ahead", "predictions", "mispredictions", "drained"], (int(v) for v in out)))

    def early_stats_counters(self) -> dict:
        """Early statistics passes issued / rounds that fell back to a regular statistics pass, since creation."""
        out = np.zeros(2, dtype=np.uint64)
        self._check(self._L.pgbart_get_early_stats_counters(self._h, out.ctypes.data_as(C.c_void_p)))
        return {"issued": int(out[0]), "fallbacks": int(out[1])}

    def profile(self) -> dict:
        """In-kernel stopwatch (SM cycles since creation), by phase; see include/pgbart.h."""
        out = np.zeros(69)
        self._check(self._L.pgbart_get_profile(self._h, out.ctypes.data_as(C.c_void_p)))
        names = ["begin", "root", "minmax", "stats", "bern", "part", "final", "done"]
        sub = ["reduce", "after_stats", "after_minmax", "finish_round", "prep_part", "finalize", "begin_update", "draw_round"]
        return {"diag": {n: dict(zip(["issue_to_last_seen_ns", "issue_to_last_arrival_ns", "issue_to_completion_ns", "longest_body_cycles"],
                                      out[57 + 4 * i:61 + 4 * i].tolist())) for i, n in enumerate(["root", "stats", "part"])},
                "control": dict(zip(["root_spec_cycles", "root_spec_n", "root_nospec_cycles", "root_nospec_n", "wait_root", "wait_stats",
                                     "wait_part", "wait_other"], out[49:57].tolist())),
                "speculative_root_passes": float(out[48]), "serial_sub_cycles": dict(zip(sub, out[32:40].tolist())), "serial_sub_calls": dict(zip(sub, out[40:48].tolist())),"pass_cycles": dict(zip(names, out[0:8].tolist())), "wait_serial_cycles": dict(zip(names, out[8:16].tolist())),
                "serial_cycles": dict(zip(names, out[16:24].tolist())), "passes": dict(zip(names, out[24:32].tolist()))}

    def predict(self, x) -> np.ndarray:
        """Sum of trees at new rows (reference TreeArrays::predict_batch_test, src/tree.rs:169-192, over the forest)."""
        x = np.asarray(x)
        if x.ndim != 2:
            raise TypeError("x must be a 2-D array")
        if x.dtype == np.float32:
            dtype = _capi.X_F32
        else:
            x = x.astype(np.float64, copy=False)
            dtype = _capi.X_F64
        if x.flags.f_contiguous:
            order = _capi.ORDER_F
        else:
            x = np.ascontiguousarray(x)
            order = _capi.ORDER_C
        out = np.empty(x.shape[0], dtype=np.float64)
        self._check(self._L.pgbart_predict(self._h, x.ctypes.data_as(C.c_void_p), dtype, order, x.shape[0], out.ctypes.data_as(C.c_void_p)))
        return out

    def variable_inclusion(self) -> np.ndarray:
        """Splits per variable in the current forest (reference BartState.variable_inclusion, src/state.rs:10)."""
        p = int(self.p)
        out = np.zeros(p, dtype=np.uint32)
        self._check(self._L.pgbart_variable_inclusion(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def debug_state(self) -> dict:
        """Where the device state machine stands (readable after a step that failed gracefully; include/pgbart.h)."""
        out = np.zeros(8)
        rc = self._L.pgbart_debug_state(self._h, out.ctypes.data_as(C.c_void_p))
        if rc:
            return {"unreadable": True}
        return dict(zip(["phase", "k", "round", "upd", "xseq", "grid_passes", "error", "t_cur"], out.tolist()))

    def block_cycles(self) -> np.ndarray:
        """[pass block][phase] cycles spent in pass bodies (only filled by -DPG_DIAG builds)."""
        g = int(self.counters()["grid_blocks"])
        out = np.zeros(g * 16)
        self._check(self._L.pgbart_debug_block_cycles(self._h, out.ctypes.data_as(C.c_void_p), out.size))
        return np.concatenate([out[:g * 8].reshape(g, 8), out[g * 8:].reshape(g, 8)], axis=1)

    def timer_start(self) -> None:
        """CUDA-event stopwatch on the sampler's stream (the stream the kernels run on)."""
        self._check(self._L.pgbart_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._check(self._L.pgbart_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.pgbart_destroy(self._h)
            self._h = None
            if getattr(self, "_pool", None) is not None:
                self._pool.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc != 0:
            msg = self._L.pgbart_last_error(self._h).decode()
            raise (ValueError if rc == 1 else RuntimeError)(msg)
