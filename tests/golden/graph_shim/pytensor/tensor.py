"""`pytensor.tensor` stand-in: elementwise Ops (NumPy perform, symbolic L_op), reductions, indexing."""
import numpy as np

from .graph.basic import Apply, Constant, TensorType, Variable
from .graph.op import Op

LN10 = float(np.log(10.0))


def as_tensor_variable(x, name=None):
    if isinstance(x, Variable):
        return x
    if hasattr(x, "detach"):
        x = x.detach().numpy()
    return Constant(np.asarray(x), name)


def dscalar(name=None):
    return TensorType(0)(name)


def dvector(name=None):
    return TensorType(1)(name)


def dmatrix(name=None):
    return TensorType(2)(name)


def tensor(name=None, ndim=0):
    return TensorType(ndim)(name)


# ---- elementwise ----------------------------------------------------------------------------------
class Elemwise(Op):
    """out = fn(*inputs) with NumPy broadcasting; `rule(inputs, out, g)` gives one symbolic gradient per
    input (None = no gradient), each then summed back to its input's shape."""

    __props__ = ("name",)
    REGISTRY = {}

    def __init__(self, name, fn, rule, out_dtype="float64"):
        self.name, self.fn, self.rule, self.out_dtype = name, fn, rule, out_dtype
        Elemwise.REGISTRY[name] = self

    def make_node(self, *inputs):
        ins = [as_tensor_variable(x) for x in inputs]
        ndim = max(i.ndim for i in ins)
        return Apply(self, ins, [TensorType(ndim, self.out_dtype)()])

    def perform(self, node, inputs, output_storage):
        with np.errstate(all="ignore"):
            output_storage[0][0] = np.asarray(self.fn(*inputs))

    def infer_shape(self, fgraph, node, input_shapes):
        return [np.broadcast_shapes(*input_shapes)]

    def L_op(self, inputs, outputs, output_grads):
        gs = self.rule(inputs, outputs[0], output_grads[0])
        return [None if g is None else sum_to_shape(g, i) for g, i in zip(gs, inputs)]


def _no_grad(n):
    return lambda ins, out, g: [None] * n


add = Elemwise("add", np.add, lambda ins, out, g: [g, g])
sub = Elemwise("sub", np.subtract, lambda ins, out, g: [g, neg(g)])
mul = Elemwise("mul", np.multiply, lambda ins, out, g: [mul(g, ins[1]), mul(g, ins[0])])
true_div = Elemwise("true_div", np.divide,
                    lambda ins, out, g: [true_div(g, ins[1]), neg(true_div(mul(g, out), ins[1]))])
neg = Elemwise("neg", np.negative, lambda ins, out, g: [neg(g)])
pow = Elemwise("pow", np.power,  # noqa: A001
               lambda ins, out, g: [None if isinstance(ins[0], Constant) else mul(g, mul(ins[1], pow(ins[0], sub(ins[1], 1.0)))),
                                    None if isinstance(ins[1], Constant) else mul(g, mul(out, log(ins[0])))])
exp = Elemwise("exp", np.exp, lambda ins, out, g: [mul(g, out)])
log = Elemwise("log", np.log, lambda ins, out, g: [true_div(g, ins[0])])
log10 = Elemwise("log10", np.log10, lambda ins, out, g: [true_div(g, mul(ins[0], LN10))])
sqrt = Elemwise("sqrt", np.sqrt, lambda ins, out, g: [true_div(g, mul(2.0, out))])
sgn = Elemwise("sgn", np.sign, _no_grad(1))
abs = Elemwise("abs", np.abs, lambda ins, out, g: [mul(g, sgn(ins[0]))])  # noqa: A001
gt = Elemwise("gt", np.greater, _no_grad(2), "bool")
lt = Elemwise("lt", np.less, _no_grad(2), "bool")
ge = Elemwise("ge", np.greater_equal, _no_grad(2), "bool")
le = Elemwise("le", np.less_equal, _no_grad(2), "bool")
and_ = Elemwise("and", np.logical_and, _no_grad(2), "bool")
switch = Elemwise("switch", np.where,
                  lambda ins, out, g: [None, switch(ins[0], g, 0.0), switch(ins[0], 0.0, g)])
# pt.clip: the gradient reaches x where lo <= x <= hi
clip = Elemwise("clip", np.clip,
                lambda ins, out, g: [switch(and_(ge(ins[0], ins[1]), le(ins[0], ins[2])), g, 0.0), None, None])
ones_like = Elemwise("ones_like", np.ones_like, _no_grad(1))
zeros_like = Elemwise("zeros_like", np.zeros_like, _no_grad(1))


def zeros(shape):
    return Constant(np.zeros(shape))


def ones(shape):
    return Constant(np.ones(shape))


def broadcast_arrays(*xs):
    xs = [as_tensor_variable(x) for x in xs]
    total = zeros_like(xs[0])
    for x in xs[1:]:
        total = add(total, zeros_like(x))
    return [add(x, total) for x in xs]


# ---- shape plumbing ---------------------------------------------------------------------------------
class SumToShape(Op):
    """g summed over the axes NumPy broadcasting added relative to `like` (the reverse of broadcasting)."""

    __props__ = ()

    def make_node(self, g, like):
        g, like = as_tensor_variable(g), as_tensor_variable(like)
        return Apply(self, [g, like], [like.type()])

    def perform(self, node, inputs, output_storage):
        g, like = inputs
        g = np.asarray(g, dtype=np.float64)
        shape = tuple(np.shape(like))
        while g.ndim > len(shape):
            g = g.sum(axis=0)
        for ax, n in enumerate(shape):
            if n == 1 and g.shape[ax] != 1:
                g = g.sum(axis=ax, keepdims=True)
        output_storage[0][0] = np.broadcast_to(g, shape).copy() if g.shape != shape else g

    def infer_shape(self, fgraph, node, input_shapes):
        return [input_shapes[1]]


def sum_to_shape(g, like):
    return SumToShape()(g, like)


class Sum(Op):
    __props__ = ("axis",)

    def __init__(self, axis=None):
        self.axis = axis

    def make_node(self, x):
        x = as_tensor_variable(x)
        return Apply(self, [x], [TensorType(0 if self.axis is None else x.ndim - 1)()])

    def perform(self, node, inputs, output_storage):
        output_storage[0][0] = np.asarray(np.sum(inputs[0], axis=self.axis))

    def infer_shape(self, fgraph, node, input_shapes):
        s = list(input_shapes[0])
        if self.axis is None:
            return [()]
        s.pop(self.axis)
        return [tuple(s)]

    def L_op(self, inputs, outputs, output_grads):
        return [ExpandLike(self.axis)(output_grads[0], inputs[0])]


class ExpandLike(Op):
    """The gradient of Sum: g put back along `axis` and broadcast to the shape of `like`."""

    __props__ = ("axis",)

    def __init__(self, axis):
        self.axis = axis

    def make_node(self, g, like):
        g, like = as_tensor_variable(g), as_tensor_variable(like)
        return Apply(self, [g, like], [like.type()])

    def perform(self, node, inputs, output_storage):
        g, like = inputs
        g = np.asarray(g, dtype=np.float64)
        if self.axis is not None:
            g = np.expand_dims(g, self.axis)
        output_storage[0][0] = np.broadcast_to(g, np.shape(like)).copy()

    def infer_shape(self, fgraph, node, input_shapes):
        return [input_shapes[1]]


def sum(x, axis=None):  # noqa: A001
    return Sum(axis)(x)


def _freeze(idx):
    if isinstance(idx, tuple):
        return tuple(_freeze(i) for i in idx)
    if isinstance(idx, slice):
        return ("slice", idx.start, idx.stop, idx.step)
    if idx is Ellipsis:
        return ("ellipsis",)
    if idx is None:
        return ("newaxis",)
    return ("int", int(idx))


def _thaw(f):
    if f and isinstance(f[0], tuple):
        return tuple(_thaw(i) for i in f)
    if f[0] == "slice":
        return slice(f[1], f[2], f[3])
    if f[0] == "ellipsis":
        return Ellipsis
    if f[0] == "newaxis":
        return None
    return f[1]


class Subtensor(Op):
    __props__ = ("frozen",)

    def __init__(self, idx):
        self.frozen = _freeze(idx if isinstance(idx, tuple) else (idx,))

    @property
    def idx(self):
        return _thaw(self.frozen)

    def make_node(self, x):
        x = as_tensor_variable(x)
        ndim = np.zeros((2,) * x.ndim)[self.idx].ndim
        return Apply(self, [x], [TensorType(ndim)()])

    def perform(self, node, inputs, output_storage):
        output_storage[0][0] = np.asarray(inputs[0][self.idx])

    def infer_shape(self, fgraph, node, input_shapes):
        return [np.zeros(input_shapes[0])[self.idx].shape]

    def L_op(self, inputs, outputs, output_grads):
        return [IncSubtensor(self.idx)(output_grads[0], inputs[0])]


class IncSubtensor(Op):
    __props__ = ("frozen",)

    def __init__(self, idx):
        self.frozen = _freeze(idx if isinstance(idx, tuple) else (idx,))

    def make_node(self, g, like):
        g, like = as_tensor_variable(g), as_tensor_variable(like)
        return Apply(self, [g, like], [like.type()])

    def perform(self, node, inputs, output_storage):
        g, like = inputs
        z = np.zeros(np.shape(like))
        z[_thaw(self.frozen)] += g
        output_storage[0][0] = z


def _r(op):
    return lambda a, b: op(b, a)


for _name, _op in (("__add__", add), ("__sub__", sub), ("__mul__", mul), ("__truediv__", true_div), ("__pow__", pow),
                   ("__gt__", gt), ("__lt__", lt), ("__ge__", ge), ("__le__", le)):
    setattr(Variable, _name, (lambda op: lambda a, b: op(a, b))(_op))
for _name, _op in (("__radd__", add), ("__rsub__", sub), ("__rmul__", mul), ("__rtruediv__", true_div), ("__rpow__", pow)):
    setattr(Variable, _name, _r(_op))
Variable.__neg__ = lambda a: neg(a)
Variable.__getitem__ = lambda a, idx: Subtensor(idx)(a)
Variable.sum = lambda a, axis=None: sum(a, axis)
Variable.__array_priority__ = 1000  # ndarray op Variable defers to the Variable
Variable.__hash__ = lambda a: id(a)
