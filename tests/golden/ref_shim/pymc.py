"""`pymc` stand-in: value replay of the free RVs + closed-form Normal likelihood (../README.md)."""
import math

import numpy as np
import torch

_STACK = []
VALUES = {}  # name -> torch tensor supplied by the generator for every free RV


class Model:
    def __init__(self, coords=None):
        self.coords = dict(coords or {})
        self.named = {}
        self.free = {}          # name -> (distribution, parameters, dims)
        self.deterministics = []
        self.observed = {}      # name -> (mu, sigma, observed)
        self.logp = 0.0         # sum of the observed log-densities

    def __enter__(self):
        _STACK.append(self)
        return self

    def __exit__(self, *exc):
        _STACK.pop()

    def __getitem__(self, name):
        return self.named[name]

    def __contains__(self, name):
        return name in self.named


def _model():
    return _STACK[-1]


def _free(dist, name, params, dims=None):
    m = _model()
    value = VALUES[name]
    if dims is None and value.ndim > 0:
        value = value[0]  # declared without a cloud dimension: a scalar RV in the reference
    m.free[name] = (dist, params, dims)
    m.named[name] = value
    return value


def _to(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64))


def Normal(name, mu=0.0, sigma=1.0, dims=None, observed=None):
    if observed is None:
        return _free("Normal", name, {"mu": mu, "sigma": sigma}, dims)
    m = _model()
    mu, sigma, obs = _to(mu), _to(sigma), _to(observed)
    # Normal log-density (PyMC's closed form)
    lp = -0.5 * ((obs - mu) / sigma) ** 2 - torch.log(sigma) - 0.5 * math.log(2.0 * math.pi)
    m.observed[name] = (mu, sigma, obs)
    m.logp = m.logp + lp.sum()
    m.named[name] = obs
    return obs


def HalfNormal(name, sigma=1.0, dims=None):
    return _free("HalfNormal", name, {"sigma": sigma}, dims)


def ChiSquared(name, nu=1, dims=None):
    return _free("ChiSquared", name, {"nu": nu}, dims)


def Beta(name, alpha=1.0, beta=1.0, dims=None):
    return _free("Beta", name, {"alpha": alpha, "beta": beta}, dims)


def Uniform(name, lower=0.0, upper=1.0, dims=None):
    return _free("Uniform", name, {"lower": lower, "upper": upper}, dims)


def TruncatedNormal(name, mu=0.0, sigma=1.0, lower=None, upper=None, dims=None):
    return _free("TruncatedNormal", name, {"mu": mu, "sigma": sigma, "lower": lower, "upper": upper}, dims)


def Deterministic(name, var, dims=None):
    m = _model()
    m.named[name] = var
    m.deterministics.append(name)
    return var


def Data(name, value, dims=None):
    m = _model()
    v = _to(value)
    m.named[name] = v
    return v
