"""Reverse-mode `grad(cost, wrt)` over the Apply graph through every Op's L_op."""
from .graph.basic import Constant, TensorType, Variable, toposort
from . import tensor as pt


class DisconnectedType(TensorType):
    def __init__(self):
        super().__init__(0)


def disconnected():
    return Variable(DisconnectedType())


def grad(cost, wrt, disconnected_inputs="zero"):
    single = isinstance(wrt, Variable)
    wrt_list = [wrt] if single else list(wrt)
    acc = {id(cost): Constant(1.0)}
    for node in reversed(toposort([cost])):
        ogs = [acc.get(id(o)) for o in node.outputs]
        if all(g is None for g in ogs):
            continue
        ogs = [g if g is not None else disconnected() for g in ogs]
        igs = node.op.L_op(node.inputs, node.outputs, ogs)
        if not isinstance(igs, (list, tuple)):
            igs = [igs]
        assert len(igs) == len(node.inputs), f"{type(node.op).__name__}.L_op returned {len(igs)} gradients for {len(node.inputs)} inputs"
        for i, g in zip(node.inputs, igs):
            if g is None or isinstance(getattr(g, "type", None), DisconnectedType) or isinstance(i, Constant):
                continue
            acc[id(i)] = g if id(i) not in acc else pt.add(acc[id(i)], g)
    out = [acc[id(w)] if id(w) in acc else pt.zeros_like(w) for w in wrt_list]
    return out[0] if single else out
