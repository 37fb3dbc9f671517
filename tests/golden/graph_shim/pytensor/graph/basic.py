"""Variable / Constant / Apply and graph evaluation through Op.perform."""
import numpy as np


class TensorType:
    def __init__(self, ndim, dtype="float64"):
        self.ndim, self.dtype = ndim, dtype

    def __call__(self, name=None):
        return Variable(self, None, None, name)

    def __eq__(self, other):
        return type(other) is type(self) and (self.ndim, self.dtype) == (other.ndim, other.dtype)

    def __hash__(self):
        return hash((type(self), self.ndim, self.dtype))


class Variable:
    def __init__(self, type_, owner=None, index=None, name=None):
        self.type, self.owner, self.index, self.name = type_, owner, index, name

    @property
    def ndim(self):
        return self.type.ndim

    @property
    def dtype(self):
        return self.type.dtype

    def eval(self, givens=None):
        givens = givens or {}
        ins = list(givens)
        return function(ins, [self])(*[givens[k] for k in ins])[0]

    # operators: filled in by pytensor.tensor (avoids a circular import)
    def __repr__(self):
        return f"<Variable {self.name or ''} ndim={self.ndim}>"


class Constant(Variable):
    def __init__(self, data, name=None):
        data = np.asarray(data)
        if data.dtype.kind in "iub" and data.dtype != bool:
            data = data.astype(np.float64)
        super().__init__(TensorType(data.ndim, "bool" if data.dtype == bool else "float64"), None, None, name)
        self.data = data


class Apply:
    def __init__(self, op, inputs, outputs):
        self.op, self.inputs, self.outputs = op, list(inputs), list(outputs)
        for i, o in enumerate(self.outputs):
            o.owner, o.index = self, i


def toposort(outputs):
    """Apply nodes needed for `outputs`, inputs before users."""
    order, seen = [], set()

    def visit(v):
        node = v.owner
        if node is None or id(node) in seen:
            return
        seen.add(id(node))
        for i in node.inputs:
            visit(i)
        order.append(node)

    import sys
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
    for o in outputs:
        visit(o)
    return order


def function(inputs, outputs):
    """Callable evaluating `outputs` from numeric `inputs` through every node's Op.perform."""
    order = toposort(outputs)

    def run(*values):
        val = {id(v): np.asarray(x, dtype=np.float64) for v, x in zip(inputs, values)}
        for node in order:
            args = []
            for i in node.inputs:
                if id(i) in val:
                    args.append(val[id(i)])
                elif isinstance(i, Constant):
                    args.append(i.data)
                else:
                    raise ValueError(f"missing input {i!r} (name {i.name})")
            storage = [[None] for _ in node.outputs]
            node.op.perform(node, args, storage)
            for o, s in zip(node.outputs, storage):
                val[id(o)] = np.asarray(s[0])
        res = []
        for o in outputs:
            res.append(val[id(o)] if id(o) in val else o.data)
        return res

    return run
