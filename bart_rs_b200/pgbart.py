"""Host-side mirror of the reference's PGBART step method (python/pymc_bart/pgbart.py:34-242).

The reference's class is a PyMC `ArrayStepShared`; PyMC is not a dependency of the hot path, so this mirror
takes the BART node's attributes directly (X, Y, m, alpha, beta, split_prior) and performs the same parameter
derivations before creating the device sampler:

    alpha_vec / cumsum split prior      pgbart.py:105-110   (passed as-is: the sampler consumes it as raw weights, Q3)
    leaf_sd = 3/sqrt(m) for 0/1 data, else std(Y)/sqrt(m)     pgbart.py:119-125
    max_depth = calculate_max_tree_depth(alpha, beta, .99)    pgbart.py:146, 213-242
    PyBartSettings(...), PySampler.init(asfortranarray(X), Y, model, settings)   pgbart.py:154-182
    astep: push likelihood parameters, step(tune), stats dict                    pgbart.py:187-201

When PyMC is importable, `as_pymc_step(...)` wraps it as an `ArrayStepShared` with the reference's name and stats.
"""
from __future__ import annotations

import math
from time import perf_counter
from typing import Optional, Sequence, Tuple

import numpy as np

from .pymc_bart import DeviceLikelihood, PyBartSettings, PySampler


def calculate_max_tree_depth(alpha: float, beta: float, probs_leaf: float) -> int:
    """Same formula and error messages as the reference (pgbart.py:213-242)."""
    if not (0 < alpha < 1):
        raise ValueError(f"alpha must be between 0 and 1. Received {alpha}")
    if not (0 < probs_leaf < 1):
        raise ValueError(f"alpha must be between 0 and 1. Received {alpha}")
    if beta <= 0:
        raise ValueError(f"beta must be greater than 0. Received {beta}")
    return int(math.pow((1 / (1 - probs_leaf)) * alpha, 1.0 / beta) - 1)


class PGBART:
    """Particle Gibbs BART sampling step on a B200 (mirror of pymc_bart.PGBART)."""

    name = "pgbart"
    generates_stats = True
    stats_dtypes_shapes = {"variable_inclusion": (object, []), "tune": (bool, []), "time": (float, [])}

    def __init__(self, X, Y, m: int = 50, alpha: float = 0.95, beta: float = 2.0,
                 split_prior: Optional[Sequence[float]] = None, split_rules: Optional[Sequence[str]] = None,
                 response: str = "constant", num_particles: int = 10, batch: Tuple[float, float] = (0.1, 0.1),
                 likelihood: Optional[DeviceLikelihood] = None, device: int = 0, seed: int = 0):
        self.X = np.asarray(X)
        self.Y = np.asarray(Y, dtype=np.float64)
        self.m = int(m)
        p = self.X.shape[1]
        if split_prior is None or len(split_prior) == 0:
            self.alpha_vec = np.ones(p)
        else:
            self.alpha_vec = np.asarray(split_prior, dtype=np.float64)
        splitting_probs = np.cumsum(self.alpha_vec)                      # pgbart.py:110
        self.split_rules = list(split_rules) if split_rules else ["ContinuousSplit"] * p
        y_unique = np.unique(self.Y)
        self.leaf_sd = np.ones(1)
        binary = y_unique.size == 2 and np.all(y_unique == [0, 1])
        if binary:
            self.leaf_sd *= 3 / self.m ** 0.5                            # pgbart.py:122-123
        else:
            self.leaf_sd *= self.Y.std() / self.m ** 0.5                 # pgbart.py:125
        if likelihood is None:
            likelihood = DeviceLikelihood("bernoulli_logit") if binary else DeviceLikelihood("gaussian", 1.0)
        self.likelihood = likelihood
        max_depth = calculate_max_tree_depth(alpha, beta, probs_leaf=0.99)
        settings = PyBartSettings(
            n_trees=self.m, n_particles=num_particles, max_depth=max_depth, alpha=alpha, beta=beta,
            sigma=float(self.leaf_sd.ravel()[0]), split_prior=splitting_probs.tolist(), split_rules=self.split_rules,
            response_rule=response, resampling_rule="systematic", batch_tune=batch[0], batch_post=batch[1], seed=seed)
        self.pg_bart = PySampler.init(x=np.asfortranarray(self.X), y=self.Y, model=self.likelihood, settings=settings,
                                      device=device)
        self.tune = True

    def astep(self, _=None):
        t0 = perf_counter()
        self.pg_bart.set_likelihood(self.likelihood)     # the reference refreshes shared likelihood inputs here (pgbart.py:190)
        sum_trees = self.pg_bart.step(self.tune)
        t1 = perf_counter()
        # pgbart.py:195-199 reports the placeholder np.array([0.0]) for variable_inclusion (the Rust side never fills
        # state.rs:10); here it is the real per-variable split count of the current forest
        stats = {"variable_inclusion": self.pg_bart.variable_inclusion().astype(np.float64), "tune": self.tune, "time": t1 - t0}
        return sum_trees, [stats]

    def predict(self, X_new):
        """Sum of trees of the current forest at new rows (reference TreeArrays::predict_batch_test, src/tree.rs:169-192)."""
        return self.pg_bart.predict(np.asarray(X_new))

    def stop_tuning(self):
        self.tune = False
