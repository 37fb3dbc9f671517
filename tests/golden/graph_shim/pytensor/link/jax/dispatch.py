from functools import singledispatch


@singledispatch
def jax_funcify(op, node=None, **kwargs):
    raise NotImplementedError(f"No jax conversion for the Op: {type(op).__name__}")
