"""The Bernoulli pass's arithmetic (csrc/pg_serial.h: pg_log_ge1, pg_bern_term) against extended precision.

The device evaluates y mu - softplus(mu), mu = o + v, as y mu - log(1 + exp(o) exp(v)) with a 64-entry table logarithm;
the reference evaluates the PyMC logp (weight.rs:26-30), the oracle y mu - (max(mu, 0) + log1p(exp(-|mu|))).  The host build
of the same source is measured here; the GPU replay tests compare the sums.
"""
import ctypes as C

import numpy as np

from tests import sim_lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_log_ge1_against_long_double():
    L = sim_lib.lib()
    rng = np.random.default_rng(0)
    u = np.concatenate([1.0 + rng.random(200_000), np.exp(rng.uniform(0.0, 600.0, 200_000)),
                        np.array([1.0, 2.0, np.nextafter(2.0, 1.0), np.nextafter(1.0, 2.0), 4.0, 1e300 * 0.99])])
    out = np.empty_like(u)
    L.pgsim_log_ge1(_p(u), _p(out), u.size)
    ref = np.log(u.astype(np.longdouble))
    err = np.abs(out.astype(np.longdouble) - ref)
    small = u <= 2.0
    assert float(err[small].max()) <= 1.2e-16
    assert float((err[~small] / np.spacing(np.abs(ref[~small]).astype(np.float64))).max()) <= 1.5


def test_bern_term_against_long_double():
    L = sim_lib.lib()
    rng = np.random.default_rng(1)
    n = 300_000
    o = np.concatenate([rng.normal(0.0, 3.0, n), rng.uniform(-299.0, 299.0, n), np.array([0.0, 301.0, -301.0, 700.0, -800.0])])
    v = np.concatenate([rng.normal(0.0, 0.3, n), rng.normal(0.0, 5.0, n), np.array([0.0, 0.1, -0.1, 1.0, -1.0])])
    y = (rng.random(o.size) < 0.5).astype(np.float64)
    out = np.empty_like(o)
    L.pgsim_bern_terms(_p(y), _p(o), _p(v), _p(out), o.size)
    mu = o.astype(np.longdouble) + v.astype(np.longdouble)
    ref = y * mu - (np.maximum(mu, 0) + np.log1p(np.exp(-np.abs(mu))))
    err = np.abs(out.astype(np.longdouble) - ref)
    # exp(o) and exp(v) carry an ulp each: a relative error of ~3 ulps in exp(o) exp(v) is an absolute error of ~7e-16 in its logarithm (bound: 1e-15); and y mu - softplus(mu) cancels down to
    # ulp(mu) for large |mu| in ANY f64 evaluation that forms softplus(mu) first (the oracle's included)
    lim = np.maximum(np.maximum(3.0 * np.spacing(np.abs(ref).astype(np.float64)), 2.5 * np.spacing(np.abs(mu).astype(np.float64))), 1e-15)
    assert bool((err <= lim).all()), float((err / lim).max())
