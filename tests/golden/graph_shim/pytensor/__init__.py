"""Symbolic PyTensor-protocol stand-in (see ../README.md)."""
from . import gradient, graph, tensor  # noqa: F401
from .graph.basic import function  # noqa: F401
