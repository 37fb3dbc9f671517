"""CPU half of the PyTensor-Op tests: the symbolic stand-in (tests/golden/graph_shim) against torch autograd,
the Ops' graph construction, and -- where the reference tree is present (build container) -- that the
committed graph fixtures are what the reference's UNMODIFIED classes build today."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "golden", "graph_shim")


def run(code):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([SHIM, ROOT, os.path.join(ROOT, "tests")]))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return res.stdout


def test_symbolic_standin_matches_torch_autograd():
    out = run("""
import numpy as np, torch, pytensor, pytensor.tensor as pt
from pytensor.gradient import grad
x, y = pt.dvector('x'), pt.dscalar('y')
f = (pt.exp(-x * y) * pt.sqrt(x) + pt.log10(x) / y + pt.switch(pt.gt(x, 1.0), x ** 2.0, 2.0 ** x) + pt.clip(x, 0.5, 1.5) * 3.0
     + (10.0 ** x)[0:1] + pt.abs(x - 1.0) + pt.log(x) * np.array([1.0, 2.0, 3.0])).sum()
gx, gy = grad(f, [x, y])
xv, yv = np.array([0.3, 1.2, 2.0]), 0.7
val, g1, g2 = pytensor.function([x, y], [f, gx, gy])(xv, yv)
xt, yt = torch.tensor(xv, requires_grad=True), torch.tensor(yv, requires_grad=True, dtype=torch.float64)
ft = (torch.exp(-xt * yt) * torch.sqrt(xt) + torch.log10(xt) / yt + torch.where(xt > 1.0, xt ** 2.0, 2.0 ** xt)
      + torch.clamp(xt, 0.5, 1.5) * 3.0 + (10.0 ** xt)[0:1] + abs(xt - 1.0) + torch.log(xt) * torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)).sum()
ft.backward()
assert abs(float(val) - float(ft.detach())) < 1e-13 and np.abs(g1 - xt.grad.numpy()).max() < 1e-13 and abs(float(g2) - float(yt.grad)) < 1e-13
print('OK')
""")
    assert "OK" in out


def test_ops_build_graphs_on_the_standin():
    """make_node / L_op / infer_shape / __props__ of every Op, without a GPU (no perform)."""
    out = run("""
import numpy as np, pytensor.tensor as pt
from pytensor.gradient import grad, DisconnectedType
from pytensor.graph.basic import toposort
from caribou_hi_b200 import ops, physics
assert ops.HAVE_PYTENSOR
tk, de, na = pt.dvector('tk'), pt.dvector('de'), pt.dvector('na')
ts = physics.calc_spin_temp(tk, de, na)                     # SpinTempOp through the drop-in physics module
assert type(ts.owner.op) is ops.SpinTempOp and ts.ndim == 1
v = np.linspace(-10, 10, 32)
vel, f2 = pt.dvector('vel'), pt.dvector('f2')
prof = physics.calc_pseudo_voigt(v, vel, f2, 0.5)           # scalar fwhm_L broadcasts
assert type(prof.owner.op) is ops.PseudoVoigtOp and prof.ndim == 2
assert prof.owner.op.infer_shape(None, prof.owner, [(3,), (3,), (3,)]) == [(32, 3)]
tb = physics.radiative_transfer(prof * 2.0, ts, 1.0, 3.77)
assert type(tb.owner.op) is ops.RadiativeTransferOp and tb.owner.op.infer_shape(None, tb.owner, [(32, 3), (3,), (3,)]) == [(32,)]
cost = (tb * np.arange(32.0)).sum()
gs = grad(cost, [tk, de, na, vel, f2])
kinds = {type(n.op).__name__ for n in toposort(gs)}
assert {'SpinTempGradOp', 'PseudoVoigtGradOp', 'RadiativeTransferGradOp'} <= kinds, kinds
assert ops.PseudoVoigtOp(v) == ops.PseudoVoigtOp(v.copy()) and ops.PseudoVoigtOp(v) != ops.PseudoVoigtOp(v + 1)
assert hash(ops.RadiativeTransferOp(3.77)) == hash(ops.RadiativeTransferOp(3.77)) and ops.SpinTempOp() == ops.SpinTempOp()
class E:  # what HISpectraOp reads from an engine while the graph is built
    n_clouds, S_e, S_a, nb_e, nb_a, has_emission, has_absorption = 3, 16, 8, 0, 0, True, True
e = E()
ins = [pt.dvector(n) for n in ops.PHYSICS_INPUTS]
mu_e, mu_a = ops.HISpectraOp(e)(*ins)
assert ops.HISpectraOp(e).infer_shape(None, None, []) == [(16,), (8,)]
g = grad(mu_a.sum(), ins)                                   # emission output disconnected -> zeros cotangent
node = [n for n in toposort(g) if type(n.op) is ops.HISpectraGradOp][0]
assert len(node.inputs) == 9 and len(node.outputs) == 7
try:
    ops.HISpectraOp(e)(*ins[:3])
    raise SystemExit('expected a TypeError')
except TypeError:
    pass
print('OK')
""")
    assert "OK" in out


def test_committed_graphs_are_what_the_reference_builds():
    if not os.path.isdir("/root/reference/caribou_hi"):
        pytest.skip("reference tree not present (GPU box)")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_graph_fixture.py"), "--check"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert res.stdout.count("same") == 15


def test_linker_registrations_numba_and_jax():
    """caribou_hi_b200.backends: every Op class gets a numba_funcify / jax_funcify registration; the Numba
    wrapper (objmode call-back into perform) is compiled with the real Numba and run on a CPU stand-in of an
    Op; the JAX wrapper is run with a minimal stand-in of jax.pure_callback."""
    out = run("""
import sys, types, numpy as np
jax = types.ModuleType('jax')          # jax is not installed: the two names the registration uses
class ShapeDtypeStruct:
    def __init__(self, shape, dtype): self.shape, self.dtype = shape, dtype
def pure_callback(fn, spec, *args):
    res = fn(*args)
    assert all(np.shape(r) == s.shape for r, s in zip(res, spec)), ([np.shape(r) for r in res], [s.shape for s in spec])
    return res
jax.ShapeDtypeStruct, jax.pure_callback = ShapeDtypeStruct, pure_callback
sys.modules['jax'] = jax
import pytensor.tensor as pt
from pytensor.link.numba.dispatch import numba_funcify
from pytensor.link.jax.dispatch import jax_funcify
from caribou_hi_b200 import backends, ops
assert sorted(backends.register_all()) == ['jax', 'numba']
for cls in backends.ALL_OPS:
    assert numba_funcify.dispatch(cls) is not numba_funcify.dispatch(object), cls
    assert jax_funcify.dispatch(cls) is not jax_funcify.dispatch(object), cls
# a CPU stand-in with SpinTempOp's / SpinTempGradOp's signatures exercises the generated wrappers
class CpuSpin(ops.SpinTempOp):
    def perform(self, node, inputs, output_storage):
        tk, de, na = inputs
        output_storage[0][0] = tk * 2.0 + de + na
class CpuSpinGrad(ops.SpinTempGradOp):
    def perform(self, node, inputs, output_storage):
        tk, de, na, g = inputs
        output_storage[0][0], output_storage[1][0], output_storage[2][0] = 2.0 * g, g, g
tk, de, na = pt.dvector('tk'), pt.dvector('de'), pt.dvector('na')
out = CpuSpin()(tk, de, na)
a, b, c = np.array([1.0, 2.0]), np.array([0.5, 0.5]), np.array([0.1, 0.2])
f = numba_funcify(CpuSpin(), node=out.owner)
assert np.allclose(f(a, b, c), a * 2 + b + c)
fj = jax_funcify(CpuSpin(), node=out.owner)
assert np.allclose(fj(a, b, c), a * 2 + b + c)
gnode = CpuSpinGrad().make_node(tk, de, na, pt.dvector('g'))
fg = numba_funcify(CpuSpinGrad(), node=gnode)
r = fg(a, b, c, np.ones(2))
assert len(r) == 3 and np.allclose(r[0], 2.0)
rj = jax_funcify(CpuSpinGrad(), node=gnode)(a, b, c, np.ones(2))
assert len(rj) == 3 and np.allclose(rj[1], 1.0)
print('OK')
""")
    assert "OK" in out
