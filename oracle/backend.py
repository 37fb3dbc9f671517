"""Tiny array-namespace shim so one restatement runs on NumPy (values) and on
torch-CPU fp64 (values + autograd gradients).  Test infrastructure only."""

import numpy as np

try:  # torch is present in this image; keep the NumPy oracle usable without it
    import torch
except Exception:  # pragma: no cover
    torch = None


class NumpyBackend:
    name = "numpy"
    pi = np.pi

    @staticmethod
    def asarray(x):
        return np.asarray(x, dtype=np.float64)

    exp = staticmethod(np.exp)
    log = staticmethod(np.log)
    log10 = staticmethod(np.log10)
    sqrt = staticmethod(np.sqrt)
    abs = staticmethod(np.abs)
    where = staticmethod(np.where)
    ones_like = staticmethod(np.ones_like)
    zeros_like = staticmethod(np.zeros_like)

    @staticmethod
    def clip(x, lo, hi):
        return np.clip(x, lo, hi)

    @staticmethod
    def cumprod(x, axis):
        return np.cumprod(x, axis=axis)

    @staticmethod
    def concatenate(xs, axis):
        return np.concatenate(xs, axis=axis)

    @staticmethod
    def sum(x, axis=None):
        return np.sum(x, axis=axis)

    @staticmethod
    def stack(xs, axis=0):
        return np.stack(xs, axis=axis)


class TorchBackend:
    name = "torch"
    pi = np.pi

    @staticmethod
    def asarray(x):
        if isinstance(x, torch.Tensor):
            return x
        return torch.as_tensor(np.asarray(x, dtype=np.float64))

    @staticmethod
    def _t(x):
        return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=torch.float64)

    @classmethod
    def exp(cls, x):
        return torch.exp(cls._t(x))

    @classmethod
    def log(cls, x):
        return torch.log(cls._t(x))

    @classmethod
    def log10(cls, x):
        return torch.log10(cls._t(x))

    @classmethod
    def sqrt(cls, x):
        return torch.sqrt(cls._t(x))

    @classmethod
    def abs(cls, x):
        return torch.abs(cls._t(x))

    @classmethod
    def where(cls, c, a, b):
        a = cls._t(a)
        b = cls._t(b)
        return torch.where(c, a, b)

    @staticmethod
    def ones_like(x):
        return torch.ones_like(x)

    @staticmethod
    def zeros_like(x):
        return torch.zeros_like(x)

    @staticmethod
    def clip(x, lo, hi):
        return torch.clamp(x, lo, hi)

    @staticmethod
    def cumprod(x, axis):
        return torch.cumprod(x, dim=axis)

    @staticmethod
    def concatenate(xs, axis):
        return torch.cat(xs, dim=axis)

    @staticmethod
    def sum(x, axis=None):
        return torch.sum(x) if axis is None else torch.sum(x, dim=axis)

    @staticmethod
    def stack(xs, axis=0):
        return torch.stack(xs, dim=axis)


def backend_of(*xs):
    """Pick the torch backend if any argument is a torch tensor, else NumPy."""
    if torch is not None:
        for x in xs:
            if isinstance(x, torch.Tensor):
                return TorchBackend
    return NumpyBackend
