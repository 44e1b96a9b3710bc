"""Synthetic inputs shared by tests and bench (SURVEY.md §8d / BASELINE.md §3)."""
from __future__ import annotations

import numpy as np


def friedman(n: int, p: int, seed: int = 0, bernoulli: bool = False):
    """X ~ U[0,1) fp32 (n,p); Friedman #1 mean; Gaussian or Bernoulli response.

    Returns (X float32 C-order, y float64 holding fp32-representable values).
    """
    assert p >= 5
    rng = np.random.default_rng(seed)
    X = rng.random((n, p), dtype=np.float32)
    x = X[:, :5].astype(np.float64)          # only the five signal columns are needed in f64 (same values as widening all of X)
    f = 10 * np.sin(np.pi * x[:, 0] * x[:, 1]) + 20 * (x[:, 2] - 0.5) ** 2 + 10 * x[:, 3] + 5 * x[:, 4]
    if bernoulli:
        pr = 1.0 / (1.0 + np.exp(-(f - f.mean())))
        y = (rng.random(n) < pr).astype(np.float64)
    else:
        y = (f + rng.standard_normal(n)).astype(np.float32).astype(np.float64)
    return X, y


def small(n: int, p: int, seed: int = 0, bernoulli: bool = False):
    """Like friedman() but valid for p < 5 (uses as many terms as there are columns)."""
    rng = np.random.default_rng(seed)
    X = rng.random((n, p), dtype=np.float32)
    x = X.astype(np.float64)
    f = np.zeros(n)
    for j in range(p):
        f += (j + 1) * x[:, j]
    if bernoulli:
        pr = 1.0 / (1.0 + np.exp(-(f - f.mean())))
        y = (rng.random(n) < pr).astype(np.float64)
    else:
        y = (f + 0.5 * rng.standard_normal(n)).astype(np.float32).astype(np.float64)
    return X, y


def pgbart_params(y: np.ndarray, m: int, p: int, alpha: float = 0.95, beta: float = 2.0):
    """Parameter derivations of PGBART.__init__ (reference python/pymc_bart/pgbart.py:105-125,146,213-242)."""
    import math

    y_unique = np.unique(y)
    if y_unique.size == 2 and np.all(y_unique == [0, 1]):
        leaf_sd = 3 / m ** 0.5
    else:
        leaf_sd = y.std() / m ** 0.5
    split_prior = np.cumsum(np.ones(p))
    max_depth = int(math.pow((1 / (1 - 0.99)) * alpha, 1.0 / beta) - 1)
    return dict(sigma=float(leaf_sd), split_prior=split_prior, max_depth=max_depth, alpha=alpha, beta=beta)


def random_cases(count: int, seed: int, n_lo: int, n_hi: int):
    """Deterministic random sampler configurations for the randomized replay-parity tests (host simulation and GPU):
    (n, p, m, P, family, lik_sigma, kwargs for cfg_for, tunes, pg_correct, grid)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        n = int(rng.integers(n_lo, n_hi + 1))
        p = int(rng.integers(2, 7))
        m = int(rng.integers(2, 7))
        P = int(rng.choice([2, 3, 4, 6, 9, 16, 20, 31, 32, 33, 48]))
        family = int(rng.choice([0, 0, 1, 2]))
        lik_sigma = float(rng.choice([1.0, 0.5, 0.05])) if family == 0 else 1.0
        alpha = float(rng.choice([0.95, 0.95, 0.9999]))
        beta = float(rng.choice([1.0, 1.5, 2.0]))
        prior = bool(rng.integers(0, 2))
        batch = (float(rng.choice([1.0, 0.5])), float(rng.choice([1.0, 0.34])))
        tunes = [bool(b) for b in rng.integers(0, 2, size=int(rng.integers(1, 4)))]
        pg_correct = bool(rng.integers(0, 4) == 0)
        grid = int(rng.integers(1, 6))
        out.append((n, p, m, P, family, lik_sigma, dict(alpha=alpha, beta=beta, max_depth=12, prior=prior, batch=batch), tunes, pg_correct, grid))
    return out
