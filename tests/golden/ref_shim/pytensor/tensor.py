"""`pytensor.tensor` stand-in: the functions caribou_hi calls, eager, torch float64."""
import numpy as np
import torch


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np.float64))


def exp(x):
    return torch.exp(_t(x))


def log(x):
    return torch.log(_t(x))


def log10(x):
    return torch.log10(_t(x))


def sqrt(x):
    return torch.sqrt(_t(x))


def abs(x):  # noqa: A001
    return torch.abs(_t(x))


def gt(a, b):
    return _t(a) > _t(b)


def lt(a, b):
    return _t(a) < _t(b)


def switch(cond, a, b):
    a, b = _t(a), _t(b)
    return torch.where(cond, a, b)


def clip(x, lo, hi):
    # pt.clip: gradient flows to x inside the interval only, like torch.clamp
    return torch.clamp(_t(x), min=lo, max=hi)


def ones_like(x):
    return torch.ones_like(_t(x))


def cumprod(x, axis=None):
    return torch.cumprod(_t(x), dim=axis)


def concatenate(xs, axis=0):
    return torch.cat([_t(x) for x in xs], dim=axis)
