"""Argument validation of the C ABI that happens before any device work (runs without a GPU):
fp32 representability of f64 inputs (a silent quantisation diverges from the reference), null handles."""
import ctypes as C

import numpy as np
import pytest

from bart_rs_b200 import DeviceLikelihood, PyBartSettings, PySampler, _capi


def settings(p):
    return PyBartSettings(4, 4, 3, 0.95, 2.0, 0.1, list(np.arange(1.0, p + 1)), ["ContinuousSplit"] * p, "constant", "systematic")


def test_f64_x_that_is_not_fp32_exact_is_refused_before_any_device_work():
    X = np.zeros((6, 2))
    X[:, 0] = 1e9 + np.arange(6)          # distinct in f64, collapses to a constant column in fp32
    X[:, 1] = np.arange(6) / 4.0
    y = np.arange(6, dtype=np.float64)
    with pytest.raises(ValueError, match=r"x\[1, 0\] = 1000000001 is not exactly representable in fp32"):
        PySampler.init(X, y, DeviceLikelihood(), settings(2))
    with pytest.raises(ValueError, match=r"x\[1, 0\]"):
        PySampler.init(np.asfortranarray(X), y, DeviceLikelihood(), settings(2))


def test_f64_y_that_is_not_fp32_exact_is_refused():
    X = (np.arange(12, dtype=np.float64).reshape(6, 2)) / 8.0
    y = np.full(6, 0.1)
    with pytest.raises(ValueError, match=r"y\[0\] = 0.1"):
        PySampler.init(X, y, DeviceLikelihood(), settings(2))


def test_rounding_can_be_accepted_explicitly():
    import torch
    X = np.random.default_rng(0).random((16, 3))       # float64, not fp32-exact
    y = X[:, 0].copy()
    if torch.cuda.is_available():
        s = PySampler.init(X, y, DeviceLikelihood(), settings(3), fp32_rounding="allow")
        s.close()
    else:                                               # passes validation, then fails loudly for lack of a device
        with pytest.raises(RuntimeError, match="no CUDA device"):
            PySampler.init(X, y, DeviceLikelihood(), settings(3), fp32_rounding="allow")
    with pytest.raises(ValueError, match="fp32_rounding"):
        PySampler.init(X, y, DeviceLikelihood(), settings(3), fp32_rounding="maybe")


def test_null_handle_is_an_error_not_a_crash():
    L = _capi.lib()
    out = np.zeros(8)
    assert L.pgbart_step(None, 1, out.ctypes.data_as(C.c_void_p)) == 1
    assert L.pgbart_step_device(None, 1) == 1
    assert L.pgbart_sync(None) == 1
    assert L.pgbart_set_likelihood(None, 0, out.ctypes.data_as(C.c_void_p), 1) == 1
    assert L.pgbart_get_counters(None, out.ctypes.data_as(C.c_void_p)) == 1
    assert L.pgbart_trace_enable(None, 4, 4) == 1
    assert L.pgbart_set_option(None, b"driver", b"persistent") == 1
    assert L.pgbart_destroy(None) == 0
    assert L.pgbart_predictions_device(None) is None


def test_pinned_pool_lends_distinct_blocks_and_recycles_them():
    """The arrays PySampler.step returns are backed by pinned blocks that return to the pool when collected."""
    import gc
    from bart_rs_b200.pymc_bart import _PinnedPool

    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]

    class FakeL:          # cudaHostAlloc needs a device; the lending logic does not
        allocs, frees = 0, 0

        def pgbart_host_alloc(self, nbytes, out):
            out._obj.value = libc.malloc(nbytes)
            FakeL.allocs += 1
            return 0

        def pgbart_host_free(self, p):
            libc.free(p)
            FakeL.frees += 1
            return 0

    pool = _PinnedPool(FakeL(), 1000, cap=2)
    a = pool.take()
    b = pool.take()
    assert a is not None and b is not None and a.ctypes.data != b.ctypes.data and a.shape == (1000,)
    assert pool.take() is None                  # cap reached: the caller falls back to an ordinary array
    a[:] = 1.0
    view = a[10:20]
    addr = a.ctypes.data
    del a
    gc.collect()
    assert pool.take() is None                  # a view keeps the block lent out
    assert view.sum() == 10.0
    del view
    gc.collect()
    c = pool.take()
    assert c is not None and c.ctypes.data == addr and FakeL.allocs == 2
    pool.close()
    del b, c
    gc.collect()
    assert FakeL.frees == 2
