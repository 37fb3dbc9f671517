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
    for x in xs:
        if isinstance(x, np.ndarray) and x.dtype == object:
            return MpmathBackend
    return NumpyBackend


class MpmathBackend:
    """Arbitrary-precision arbiter (object arrays of mpmath.mpf): used only by tests to
    decide, when the CUDA path and the fp64 oracle differ at the 1e-10 level, which of
    the two is closer to the exact value."""

    name = "mpmath"

    @staticmethod
    def _mp():
        import mpmath

        return mpmath

    @classmethod
    def asarray(cls, x):
        mp = cls._mp()
        a = np.asarray(x)
        if a.dtype == object:
            return a
        return np.vectorize(lambda v: mp.mpf(float(v)), otypes=[object])(a)

    @classmethod
    def _u(cls, fn):
        return np.frompyfunc(fn, 1, 1)

    @classmethod
    def exp(cls, x):
        return cls._u(cls._mp().exp)(x)

    @classmethod
    def log(cls, x):
        return cls._u(cls._mp().log)(x)

    @classmethod
    def log10(cls, x):
        return cls._u(cls._mp().log10)(x)

    @classmethod
    def sqrt(cls, x):
        return cls._u(cls._mp().sqrt)(x)

    @classmethod
    def abs(cls, x):
        return cls._u(abs)(x)

    @staticmethod
    def where(c, a, b):
        return np.where(c, a, b)

    @classmethod
    def ones_like(cls, x):
        return cls.asarray(np.ones(np.shape(x)))

    @classmethod
    def zeros_like(cls, x):
        return cls.asarray(np.zeros(np.shape(x)))

    @staticmethod
    def clip(x, lo, hi):
        return np.where(x < lo, lo, np.where(x > hi, hi, x))

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
