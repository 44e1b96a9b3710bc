"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/pgbart.h
declares, and fails loudly (no CPU fallback) when asked to create a sampler without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "pgbart.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pgbart_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_reference_surface():
    names = declared_functions()
    for must in ("pgbart_create", "pgbart_step", "pgbart_set_likelihood", "pgbart_destroy", "pgbart_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from bart_rs_b200 import _capi
    L = _capi.lib()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/pgbart.h but not exported"
    assert sorted(_capi.EXPORTS) == names, "bart_rs_b200/_capi.py EXPORTS is out of sync with include/pgbart.h"


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible")
    from bart_rs_b200 import DeviceLikelihood, PyBartSettings, PySampler
    X = np.random.default_rng(0).random((16, 3))
    y = X[:, 0]
    st = PyBartSettings(4, 4, 3, 0.95, 2.0, 0.1, [1.0, 2.0, 3.0], ["ContinuousSplit"] * 3, "constant", "systematic")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        PySampler.init(X, y, DeviceLikelihood("gaussian", 1.0), st, fp32_rounding="allow")


def test_reference_error_behaviour_of_settings_and_init():
    from bart_rs_b200 import DeviceLikelihood, PyBartSettings, PySampler
    # PyBartSettings keeps the reference's 13 fields and defaults (src/lib.rs:54-104)
    st = PyBartSettings(n_trees=5, n_particles=3, max_depth=4, alpha=0.9, beta=1.5, sigma=0.3, split_prior=[1.0],
                        split_rules=["ContinuousSplit"], response_rule="constant", resampling_rule="systematic")
    assert (st.batch_tune, st.batch_post, st.seed) == (0.1, 0.1, 0)
    with pytest.raises(OverflowError):
        PyBartSettings(5, 3, 300, 0.9, 1.5, 0.3, [], [], "constant", "systematic")   # max_depth is u8
    X = np.zeros((4, 1))
    y = np.zeros(4)
    # the reference's `model` is a host callback address: rejected, never silently replaced
    with pytest.raises(ValueError, match="no CPU fallback"):
        PySampler.init(X, y, 140737488355328, st)
    # unknown split rule: ValueError with the reference's message (src/lib.rs:129-132, splitting.rs:117-120)
    bad = PyBartSettings(5, 3, 4, 0.9, 1.5, 0.3, [1.0], ["NoSuchRule"], "constant", "systematic")
    with pytest.raises(ValueError, match="Unknown split rule: 'NoSuchRule'"):
        PySampler.init(X, y, DeviceLikelihood(), bad)
    with pytest.raises(ValueError, match="not available on the device path"):
        DeviceLikelihood("poisson").code()
