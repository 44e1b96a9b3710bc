"""Host-side mirror of the reference's PGBART step method (python/pymc_bart/pgbart.py:34-242).

The reference's class is a PyMC `ArrayStepShared`; PyMC is not a dependency of the hot path, so `PGBART` here takes
the BART node's attributes directly (X, Y, m, alpha, beta, split_prior) and performs the same parameter derivations
before creating the device sampler:

    alpha_vec / cumsum split prior      pgbart.py:105-110   (passed as-is: the sampler consumes it as raw weights, Q3)
    leaf_sd = 3/sqrt(m) for 0/1 data, else std(Y)/sqrt(m)     pgbart.py:119-125
    max_depth = calculate_max_tree_depth(alpha, beta, .99)    pgbart.py:146, 213-242
    PyBartSettings(...), PySampler.init(asfortranarray(X), Y, model, settings)   pgbart.py:154-182
    astep: push likelihood parameters, step(tune), stats dict                    pgbart.py:187-201

The likelihood hand-off.  The reference compiles `model.datalogp` into a host callback
(compile_pymc.py:123-191, 256-307) and refreshes its shared inputs before every step (pgbart.py:190).  A host callback
cannot run inside the device-resident sweep, so `likelihood_from_model` recognises the model's observed variable
instead and maps it to a device family plus a *source* for its parameters:

    Normal(mu=<BART>, sigma=s)            -> DeviceLikelihood("gaussian", sigma) with sigma re-evaluated before each step
    Bernoulli(logit_p=<BART>)             -> DeviceLikelihood("bernoulli_logit")     (PyMC stores p = sigmoid(logit_p))
    anything else                         -> NotImplementedError (no CPU fallback)

`as_pymc_step` builds the `ArrayStepShared` subclass with the reference's name, constructor and stats when PyMC is
importable (it is not in this image; the recognition logic is unit-tested against duck-typed graph stubs,
tests/test_pymc_handoff.py).
"""
from __future__ import annotations

import math
import warnings
from time import perf_counter
from typing import Callable, Optional, Sequence, Tuple

import numpy as np

from .pymc_bart import DeviceLikelihood, PyBartSettings, PySampler, not_fp32_exact


def calculate_max_tree_depth(alpha: float, beta: float, probs_leaf: float) -> int:
    """Same formula and error messages as the reference (pgbart.py:213-242)."""
    if not (0 < alpha < 1):
        raise ValueError(f"alpha must be between 0 and 1. Received {alpha}")
    if not (0 < probs_leaf < 1):
        raise ValueError(f"alpha must be between 0 and 1. Received {alpha}")
    if beta <= 0:
        raise ValueError(f"beta must be greater than 0. Received {beta}")
    return int(math.pow((1 / (1 - probs_leaf)) * alpha, 1.0 / beta) - 1)


# ---------------------------------------------------------------------------------------------------------
# Likelihood recognition (reference: compile_pymc.py:123-191 compiles model.datalogp wholesale)

def _op_name(var) -> str:
    """Name of the random-variable op behind a (duck-typed) PyTensor variable: 'normal', 'bernoulli', ..."""
    owner = getattr(var, "owner", None)
    op = getattr(owner, "op", None)
    name = getattr(op, "name", None)
    if name is None and op is not None:
        name = type(op).__name__
    return str(name or "").lower()


def _dist_params(var):
    """Distribution parameters of an RV node.  PyMC 5: op.dist_params(node); older layouts: inputs after (rng, size)."""
    node = var.owner
    op = node.op
    if hasattr(op, "dist_params"):
        return list(op.dist_params(node))
    return list(node.inputs[2:])          # RandomVariable node layout: (rng, size, *dist_params)


def _is_bart(v, bart_rv) -> bool:
    """v is the BART variable itself (or an alias PyMC wraps around it: identity / specify_shape / deterministic)."""
    seen = 0
    while v is not None and seen < 8:
        if v is bart_rv:
            return True
        owner = getattr(v, "owner", None)
        if owner is None or len(getattr(owner, "inputs", ())) != 1:
            return False
        opn = type(owner.op).__name__.lower()
        if not any(k in opn for k in ("identity", "specifyshape", "viewop", "deepcopy")):
            return False
        v = owner.inputs[0]
        seen += 1
    return False


def _sigmoid_arg(v):
    """If v = sigmoid(u) (what pm.Bernoulli(logit_p=u) builds), return u, else None."""
    owner = getattr(v, "owner", None)
    if owner is None or len(owner.inputs) != 1:
        return None
    scalar_op = getattr(owner.op, "scalar_op", None)
    name = (type(scalar_op).__name__ if scalar_op is not None else type(owner.op).__name__).lower()
    return owner.inputs[0] if "sigmoid" in name or "expit" in name else None


def _constant_value(v):
    """Python float of a constant scalar parameter, or None if it depends on the graph."""
    if isinstance(v, (int, float, np.floating, np.integer)):
        return float(v)
    data = getattr(v, "data", None)
    if data is not None and getattr(v, "owner", None) is None:
        a = np.asarray(data, dtype=np.float64)
        if a.size == 1:
            return float(a.ravel()[0])
        if a.size > 1 and np.all(a == a.ravel()[0]):
            return float(a.ravel()[0])
    return None


def likelihood_from_model(model, bart_rv, compile_scalar: Optional[Callable] = None):
    """Map the observed variable that consumes `bart_rv` to (DeviceLikelihood, refresh).

    `refresh()` returns the likelihood with its parameters at the model's CURRENT point (the role of
    CompiledPyMCModel.update_shared_arrays, pgbart.py:190): for Normal(mu=BART, sigma=s) it re-evaluates s.
    `compile_scalar(expr) -> callable` turns a parameter expression into a zero-argument function of the sampler's
    shared variables (as_pymc_step passes one built on pytensor.function with the step's shared replacements).
    Raises NotImplementedError for every observed distribution / link the device path does not evaluate: there is no
    CPU fallback, and a silently wrong likelihood would bias the particle weights.
    """
    observed = list(getattr(model, "observed_RVs", []))
    users = []
    for rv in observed:
        name = _op_name(rv)
        try:
            params = _dist_params(rv)
        except Exception:
            params = []
        if name.startswith("normal") and len(params) >= 2:
            mu, sigma = params[0], params[1]
            if _is_bart(mu, bart_rv):
                users.append(("gaussian", sigma))
                continue
        if name.startswith("bernoulli") and len(params) >= 1:
            u = _sigmoid_arg(params[0])
            if u is not None and _is_bart(u, bart_rv):
                users.append(("bernoulli_logit", None))
                continue
        if any(_is_bart(q, bart_rv) or (_sigmoid_arg(q) is not None and _is_bart(_sigmoid_arg(q), bart_rv)) for q in params):
            raise NotImplementedError(
                f"observed distribution {name!r} of the BART variable is not evaluated on the device path "
                "(supported: Normal(mu=BART, sigma=...) and Bernoulli(logit_p=BART)); there is no CPU fallback")
    if len(users) != 1:
        raise NotImplementedError(
            "the device path needs exactly one observed Normal(mu=BART, sigma=...) or Bernoulli(logit_p=BART) whose mean is "
            f"the BART variable itself (found {len(users)} among {len(observed)} observed variables); transformed means, "
            "other links and other families would need the reference's host callback, which cannot run in the device sweep")
    family, sigma = users[0]
    if family == "bernoulli_logit":
        lik = DeviceLikelihood("bernoulli_logit")
        return lik, (lambda: lik)
    const = _constant_value(sigma)
    if const is not None:
        lik = DeviceLikelihood("gaussian", const)
        return lik, (lambda: lik)
    if compile_scalar is None:
        raise NotImplementedError("sigma of the observed Normal depends on other model variables: pass compile_scalar "
                                  "(as_pymc_step does) so that it can be re-evaluated before every step")
    fn = compile_scalar(sigma)
    lik = DeviceLikelihood("gaussian", float(np.asarray(fn()).ravel()[0]))

    def refresh():
        v = float(np.asarray(fn()).ravel()[0])
        if not (v > 0.0) or not np.isfinite(v):
            raise ValueError(f"observed Normal sigma evaluated to {v}")
        lik.sigma = v
        return lik
    return lik, refresh


class PGBART:
    """Particle Gibbs BART sampling step on a B200 (mirror of pymc_bart.PGBART without the PyMC base class).

    likelihood: DeviceLikelihood for the observed variable.  For 0/1 data it defaults to Bernoulli-logit; for anything
    else it MUST describe the model's noise (the reference takes it from the PyMC model): when omitted a Gaussian with
    sigma = 1 is assumed and a warning is raised.  sigma_source: optional zero-argument callable returning the current
    sigma; astep() calls it before every step (what update_shared_arrays does in the reference, pgbart.py:190).  Without
    it the caller must keep `self.likelihood.sigma` current.
    split_rules: "ContinuousSplit" / "OneHotSplit" per column (reference pgbart.py:146-152 -> splitting.rs); a OneHotSplit
    column holds integer codes spanning fewer than 64 values and runs the sampler on the generic control code.
    weight_mode: "reference" (default: the reference's arithmetic, stale-weight rule and fresh-stump reference particle
    included) | "pg_correct" (conditional SMC with log-weights carried by the particles; not the reference's behaviour).
    fp32_rounding: "warn" (default) | "allow" | "error" — the device stores X and Y as fp32; float64 data that is not
    fp32-representable is rounded with a warning (or silently / refused).
    """

    name = "pgbart"
    generates_stats = True
    stats_dtypes_shapes = {"variable_inclusion": (object, []), "tune": (bool, []), "time": (float, [])}

    def __init__(self, X, Y, m: int = 50, alpha: float = 0.95, beta: float = 2.0,
                 split_prior: Optional[Sequence[float]] = None, split_rules: Optional[Sequence[str]] = None,
                 response: str = "constant", num_particles: int = 10, batch: Tuple[float, float] = (0.1, 0.1),
                 likelihood: Optional[DeviceLikelihood] = None, device: int = 0, seed: int = 0,
                 sigma_source: Optional[Callable[[], float]] = None, fp32_rounding: str = "warn",
                 weight_mode: str = "reference"):
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
            if binary:
                likelihood = DeviceLikelihood("bernoulli_logit")
            else:
                warnings.warn("PGBART: no likelihood given for non-binary Y; assuming Normal(mu=BART, sigma=1). The reference "
                              "takes the likelihood from the PyMC model — pass likelihood=DeviceLikelihood('gaussian', sigma) "
                              "and keep sigma current (sigma_source=...) or the particle weights use the wrong noise scale.",
                              UserWarning, stacklevel=2)
                likelihood = DeviceLikelihood("gaussian", 1.0)
        self.likelihood = likelihood
        self.sigma_source = sigma_source
        if fp32_rounding not in ("warn", "allow", "error"):
            raise ValueError("fp32_rounding must be 'warn', 'allow' or 'error'")
        if fp32_rounding == "warn" and (not_fp32_exact(self.X) or not_fp32_exact(self.Y)):
            warnings.warn("PGBART: X / Y hold float64 values that are not exactly representable in fp32; the device stores "
                          "them rounded to fp32 (the reference computes on the float64 values). Pass fp32_rounding='allow' "
                          "to silence this or 'error' to refuse.", UserWarning, stacklevel=2)
        max_depth = calculate_max_tree_depth(alpha, beta, probs_leaf=0.99)
        settings = PyBartSettings(
            n_trees=self.m, n_particles=num_particles, max_depth=max_depth, alpha=alpha, beta=beta,
            sigma=float(self.leaf_sd.ravel()[0]), split_prior=splitting_probs.tolist(), split_rules=self.split_rules,
            response_rule=response, resampling_rule="systematic", batch_tune=batch[0], batch_post=batch[1], seed=seed)
        self.pg_bart = PySampler.init(x=np.asfortranarray(self.X), y=self.Y, model=self.likelihood, settings=settings,
                                      device=device, fp32_rounding="error" if fp32_rounding == "error" else "allow")
        if weight_mode != "reference":      # "pg_correct": NOT the reference's sampler (include/pgbart.h, option weight_mode)
            self.pg_bart.set_option("weight_mode", weight_mode)
        self.tune = True

    def astep(self, _=None):
        t0 = perf_counter()
        if self.sigma_source is not None and self.likelihood.family == "gaussian":
            self.likelihood.sigma = float(self.sigma_source())
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


def as_pymc_step():
    """Return the `PGBART(ArrayStepShared)` class for `pm.sample(step=...)` — the reference's constructor
    (vars, num_particles, batch, model, initial_point, compile_kwargs; pgbart.py:61-69), name and stats.

    Needs PyMC + PyTensor (absent from this image: ImportError).  The BART variable's op must expose X, Y, m, alpha,
    beta, split_prior, split_rules, response like pymc_bart.BARTRV (bart.py)."""
    try:
        import pytensor
        from pymc.model import modelcontext
        from pymc.pytensorf import inputvars, make_shared_replacements
        from pymc.step_methods.arraystep import ArrayStepShared
        from pymc.step_methods.compound import Competence
    except ImportError as e:      # pragma: no cover - PyMC is not in this image
        raise ImportError("as_pymc_step needs pymc and pytensor") from e

    class PGBARTStep(ArrayStepShared):   # pragma: no cover - exercised only where PyMC is installed
        name = "pgbart"
        default_blocked = False
        generates_stats = True
        stats_dtypes_shapes = {"variable_inclusion": (object, []), "tune": (bool, []), "time": (float, [])}

        def __init__(self, vars=None, num_particles=10, batch=(0.1, 0.1), model=None, initial_point=None,
                     compile_kwargs=None, device=0, seed=0, fp32_rounding="warn"):
            model = modelcontext(model)
            if initial_point is None:
                initial_point = model.initial_point()
            if vars is None:
                vars = model.value_vars
            else:
                vars = inputvars([model.rvs_to_values.get(v, v) for v in vars])
            value_bart = vars[0]
            bart_rv = model.values_to_rvs[value_bart]
            bart = bart_rv.owner.op
            X = bart.X.eval() if hasattr(bart.X, "eval") else bart.X
            Y = bart.Y.eval() if hasattr(bart.Y, "eval") else bart.Y
            shared = make_shared_replacements(initial_point, vars, model)

            def compile_scalar(expr):
                fn = pytensor.function([], pytensor.clone_replace(expr, replace=shared), on_unused_input="ignore")
                return fn

            lik, self._refresh = likelihood_from_model(model, bart_rv, compile_scalar)
            rules = list(bart.split_rules.values()) if isinstance(bart.split_rules, dict) else list(bart.split_rules or [])
            self._impl = PGBART(X, Y, m=bart.m, alpha=bart.alpha, beta=bart.beta,
                                split_prior=None if np.size(bart.split_prior) == 0 else bart.split_prior,
                                split_rules=rules or None, response=bart.response, num_particles=num_particles,
                                batch=batch, likelihood=lik, device=device, seed=seed, fp32_rounding=fp32_rounding)
            self.tune = True
            super().__init__(vars, shared)

        def astep(self, _):
            self._impl.likelihood = self._refresh()      # shared variables were updated by ArrayStepShared.step
            self._impl.tune = self.tune
            return self._impl.astep()

        @staticmethod
        def competence(var, has_grad):
            dist = getattr(var.owner, "op", None)
            return Competence.IDEAL if type(dist).__name__ == "BARTRV" else Competence.INCOMPATIBLE

    return PGBARTStep
