"""`pymc` stand-in, symbolic: free RVs are named graph inputs, the observed Normal adds its closed-form
log-density (PyMC's, third-party) to `model.logp` (see README.md)."""
import math

import numpy as np
import pytensor.tensor as pt

_STACK = []


class Model:
    def __init__(self, coords=None):
        self.coords = dict(coords or {})
        self.named = {}
        self.free = {}            # name -> symbolic input Variable
        self.free_info = {}       # name -> (distribution, dims)
        self.deterministics = []
        self.observed = {}
        self.logp = None

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


def _free(dist, name, dims):
    m = _model()
    v = pt.tensor(name, ndim=0 if dims is None else 1)
    m.free[name] = v
    m.free_info[name] = (dist, dims)
    m.named[name] = v
    return v


def Normal(name, mu=0.0, sigma=1.0, dims=None, observed=None):
    if observed is None:
        return _free("Normal", name, dims)
    m = _model()
    obs = pt.as_tensor_variable(observed)
    sigma = pt.as_tensor_variable(sigma)
    lp = -0.5 * ((obs - mu) / sigma) ** 2.0 - pt.log(sigma) - 0.5 * math.log(2.0 * math.pi)
    m.observed[name] = mu
    m.logp = lp.sum() if m.logp is None else m.logp + lp.sum()
    m.named[name] = obs
    return obs


def HalfNormal(name, sigma=1.0, dims=None):
    return _free("HalfNormal", name, dims)


def ChiSquared(name, nu=1, dims=None):
    return _free("ChiSquared", name, dims)


def Beta(name, alpha=1.0, beta=1.0, dims=None):
    return _free("Beta", name, dims)


def Uniform(name, lower=0.0, upper=1.0, dims=None):
    return _free("Uniform", name, dims)


def TruncatedNormal(name, mu=0.0, sigma=1.0, lower=None, upper=None, dims=None):
    return _free("TruncatedNormal", name, dims)


def Deterministic(name, var, dims=None):
    m = _model()
    m.named[name] = var
    m.deterministics.append(name)
    return var


def Data(name, value, dims=None):
    m = _model()
    v = pt.as_tensor_variable(np.asarray(value, dtype=np.float64), name)
    m.named[name] = v
    return v
