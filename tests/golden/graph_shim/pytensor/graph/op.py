"""Op base class: the protocol caribou_hi_b200/ops.py is written against."""


class Op:
    __props__ = ()

    def make_node(self, *inputs):
        raise NotImplementedError

    def perform(self, node, inputs, output_storage):
        raise NotImplementedError

    def L_op(self, inputs, outputs, output_grads):
        return self.grad(inputs, output_grads)

    def grad(self, inputs, output_grads):
        raise NotImplementedError(f"{type(self).__name__} has no gradient")

    def infer_shape(self, fgraph, node, input_shapes):
        raise NotImplementedError

    def __call__(self, *inputs):
        node = self.make_node(*inputs)
        return node.outputs[0] if len(node.outputs) == 1 else list(node.outputs)

    def _props(self):
        return tuple(getattr(self, p) for p in self.__props__)

    def __eq__(self, other):
        return type(self) is type(other) and self._props() == other._props()

    def __hash__(self):
        return hash((type(self), self._props()))
