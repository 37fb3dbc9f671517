from functools import singledispatch


@singledispatch
def numba_funcify(op, node=None, **kwargs):
    raise NotImplementedError(f"No numba conversion for the Op: {type(op).__name__}")
