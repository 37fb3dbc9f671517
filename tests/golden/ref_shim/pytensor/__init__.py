"""Eager torch stand-in for the part of PyTensor the reference imports (see ../README.md)."""
from . import tensor  # noqa: F401
