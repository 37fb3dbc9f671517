"""Registrations of the CUDA Ops with PyTensor's alternative linkers: the reference's
environment lists nutpie / numba and jax / blackjax / numpyro samplers (environment.yml:11-21), which compile
the model graph through ``numba_funcify`` / ``jax_funcify`` and fail on an Op they have no registration for.

Every Op here is a host call into the CUDA library, so the registrations are callbacks: ``numba.objmode`` for
the Numba linker, ``jax.pure_callback`` for the JAX linker (gradients need nothing extra -- ``L_op`` already
builds graphs of the gradient Ops, which are registered the same way).  Written against PyTensor's documented
dispatch API (``pytensor.link.numba.dispatch.numba_funcify`` and ``pytensor.link.jax.dispatch.jax_funcify``
are ``functools.singledispatch`` functions called as ``funcify(op, node=node, **kwargs)`` and returning the
callable that computes the node's outputs from its inputs); none of PyTensor, JAX or nutpie is installable in
this image, so the test-suite runs the registrations on the protocol stand-in of tests/golden/graph_shim
(the Numba path with the real Numba, which is installed).
"""

import numpy as np

from . import ops

FORWARD_OPS = (ops.SpinTempOp, ops.PseudoVoigtOp, ops.RadiativeTransferOp, ops.HISpectraOp, ops.HILogLikeOp)
GRAD_OPS = (ops.SpinTempGradOp, ops.PseudoVoigtGradOp, ops.RadiativeTransferGradOp, ops.HISpectraGradOp, ops.HILogLikeGradOp)
ALL_OPS = FORWARD_OPS + GRAD_OPS


def _perform(op, node, n_out):
    """Python callable ``(*inputs) -> tuple of float64 arrays`` running ``op.perform``."""

    def call(*inputs):
        storage = [[None] for _ in range(n_out)]
        op.perform(node, [np.asarray(x, dtype=np.float64) for x in inputs], storage)
        return tuple(np.ascontiguousarray(s[0], dtype=np.float64) for s in storage)

    return call


def _out_ndims(node):
    return [o.type.ndim for o in node.outputs]


def register_numba():
    """``numba_funcify`` registrations: the node becomes an ``objmode`` call-back into ``perform``."""
    import numba
    from pytensor.link.numba.dispatch import numba_funcify

    def funcify(op, node=None, **kwargs):
        ndims = _out_ndims(node)
        call = _perform(op, node, len(ndims))
        # objmode wants the output types spelled out and a fixed arity, so the wrapper is generated:
        #   def wrapper(a0, a1):
        #       with numba.objmode(r0='float64[:]', r1='float64[:, :]'):
        #           r0, r1 = call(a0, a1)
        #       return r0, r1
        args = ", ".join(f"a{i}" for i in range(len(node.inputs)))
        outs = [f"r{i}" for i in range(len(ndims))]
        types = ", ".join(f"{r}='float64[{', '.join([':'] * nd)}]'" if nd else f"{r}='float64'" for r, nd in zip(outs, ndims))
        unpack = ", ".join(outs) + ("," if len(outs) == 1 else "")
        scalars = "".join(f"\n        {r} = float({r})" for r, nd in zip(outs, ndims) if nd == 0)
        src = (f"def wrapper({args}):\n    with numba.objmode({types}):\n        {unpack} = call({args}){scalars}\n"
               f"    return {', '.join(outs)}\n")
        env = {"numba": numba, "call": call}
        exec(src, env)
        return numba.njit(env["wrapper"])

    for cls in ALL_OPS:
        numba_funcify.register(cls)(funcify)
    return funcify


def register_jax():
    """``jax_funcify`` registrations: ``jax.pure_callback`` into ``perform`` with the output shapes from
    ``infer_shape`` (forward Ops) or the inputs (gradient Ops return one array per differentiated input)."""
    import jax
    from pytensor.link.jax.dispatch import jax_funcify

    def funcify(op, node=None, **kwargs):
        n_out = len(node.outputs)
        call = _perform(op, node, n_out)
        is_grad = isinstance(op, GRAD_OPS)

        def fn(*inputs):
            shapes_in = [tuple(np.shape(x)) for x in inputs]
            if is_grad:
                shapes = shapes_in[:n_out]
            else:
                shapes = [tuple(s) for s in op.infer_shape(None, node, shapes_in)]
            spec = tuple(jax.ShapeDtypeStruct(s, np.float64) for s in shapes)
            res = jax.pure_callback(call, spec, *inputs)
            return res[0] if n_out == 1 else tuple(res)

        return fn

    for cls in ALL_OPS:
        jax_funcify.register(cls)(funcify)
    return funcify


def register_all():
    """Register with every linker that can be imported; returns the names of those that were."""
    done = []
    for name, fn in (("numba", register_numba), ("jax", register_jax)):
        try:
            fn()
            done.append(name)
        except ImportError:
            pass
    return done
