"""the likelihood hand-off that replaces the reference's compiled `model.datalogp` callback
(python/pymc_bart/compile_pymc.py:123-191, pgbart.py:128-201).  PyMC / PyTensor are not installed in this image, so the
recognition logic is exercised on duck-typed stand-ins for the graph objects it reads: `var.owner.op.name`,
`op.dist_params(node)`, `Elemwise.scalar_op`, constants with `.data`, and `model.observed_RVs`."""
import warnings

import numpy as np
import pytest

from bart_rs_b200.pgbart import likelihood_from_model, calculate_max_tree_depth
from bart_rs_b200.pymc_bart import not_fp32_exact


class Op:
    def __init__(self, name=None, scalar_op=None):
        if name is not None:
            self.name = name
        if scalar_op is not None:
            self.scalar_op = scalar_op

    def dist_params(self, node):
        return node.inputs[2:]


class Sigmoid:          # pytensor.scalar.math.Sigmoid
    pass


class Exp:
    pass


class Elemwise(Op):
    pass


class Node:
    def __init__(self, op, inputs):
        self.op, self.inputs = op, list(inputs)


class Var:
    def __init__(self, owner=None, data=None, name=None):
        self.owner, self.name = owner, name
        if data is not None:
            self.data = data


def rv(opname, *params):
    return Var(Node(Op(opname), [Var(name="rng"), Var(name="size"), *params]))


class Model:
    def __init__(self, *observed):
        self.observed_RVs = list(observed)


def test_normal_with_constant_sigma():
    mu = Var(name="mu_bart")
    lik, refresh = likelihood_from_model(Model(rv("normal", mu, Var(data=np.array(0.7)))), mu)
    assert (lik.family, lik.sigma) == ("gaussian", 0.7)
    assert refresh() is lik


def test_normal_sigma_is_refreshed_every_step():
    mu, sigma = Var(name="mu_bart"), Var(owner=Node(Elemwise(scalar_op=Exp()), [Var(name="sigma_log__")]))
    current = {"v": 1.5}
    seen = []

    def compile_scalar(expr):
        seen.append(expr)
        return lambda: np.array(current["v"])

    lik, refresh = likelihood_from_model(Model(rv("normal", mu, sigma)), mu, compile_scalar)
    assert seen == [sigma] and lik.sigma == 1.5
    current["v"] = 0.25
    assert refresh().sigma == 0.25
    current["v"] = -1.0
    with pytest.raises(ValueError):
        refresh()
    with pytest.raises(NotImplementedError, match="compile_scalar"):
        likelihood_from_model(Model(rv("normal", mu, sigma)), mu)


def test_bernoulli_logit_link():
    mu = Var(name="mu_bart")
    p = Var(owner=Node(Elemwise(scalar_op=Sigmoid()), [mu]))
    lik, _ = likelihood_from_model(Model(rv("bernoulli", p)), mu)
    assert lik.family == "bernoulli_logit"


def test_everything_else_raises_like_a_missing_callback():
    mu = Var(name="mu_bart")
    other = Var(name="beta")
    with pytest.raises(NotImplementedError, match="poisson"):
        likelihood_from_model(Model(rv("poisson", mu)), mu)
    with pytest.raises(NotImplementedError):                       # Bernoulli(p=BART): probability link, not logit
        likelihood_from_model(Model(rv("bernoulli", mu)), mu)
    with pytest.raises(NotImplementedError, match="exactly one"):  # the observed variable does not use the BART mean
        likelihood_from_model(Model(rv("normal", other, Var(data=1.0))), mu)
    with pytest.raises(NotImplementedError, match="exactly one"):  # two observed variables on one BART mean
        likelihood_from_model(Model(rv("normal", mu, Var(data=1.0)), rv("normal", mu, Var(data=2.0))), mu)


def test_as_pymc_step_needs_pymc():
    from bart_rs_b200.pgbart import as_pymc_step
    try:
        import pymc  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="pymc"):
            as_pymc_step()
    else:
        assert as_pymc_step().name == "pgbart"


def test_parameter_derivations_match_the_reference():
    assert calculate_max_tree_depth(0.95, 2.0, 0.99) == 8          # pgbart.py:146,241 at the defaults
    with pytest.raises(ValueError, match="alpha must be between 0 and 1"):
        calculate_max_tree_depth(1.5, 2.0, 0.99)
    with pytest.raises(ValueError, match="beta must be greater than 0"):
        calculate_max_tree_depth(0.95, 0.0, 0.99)


def test_not_fp32_exact():
    assert not not_fp32_exact(np.array([0.5, 1.0, np.nan, np.inf, 3.0e9 - 3.0e9 % 256]))
    assert not_fp32_exact(np.array([0.1]))
    assert not not_fp32_exact(np.array([0.1], dtype=np.float32))
    assert not_fp32_exact(np.array([1e9 + 1.0]))
