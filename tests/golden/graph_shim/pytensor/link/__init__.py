"""Linker dispatch points of the stand-in: `numba_funcify` / `jax_funcify` are singledispatch functions keyed by
Op class, called as funcify(op, node=node) and returning the callable that computes the node (PyTensor's
documented extension API for custom Ops)."""
