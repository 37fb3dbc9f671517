"""PyTensor ``Op`` shims: ``perform`` (forward) and ``L_op`` (gradient) call the CUDA
library through :class:`caribou_hi_b200.engine.Engine`, so a ``bayes_spec``/PyMC model
built from caribou_hi can use the kernels unchanged (INTEGRATION.md shows the
three-line patch of ``add_likelihood``).

PyTensor is a third-party dependency of the reference that is NOT installable in this
image (no network, not in the wheelhouse).  The Ops are therefore written against the
documented protocol -- ``make_node(*inputs) -> Apply``, ``perform(node, inputs,
output_storage)`` writing ``output_storage[i][0]``, ``L_op(inputs, outputs,
output_grads)``, ``infer_shape`` and ``__props__`` -- and every one of those methods is
executed by the test-suite on a symbolic stand-in of that protocol
(``tests/golden/graph_shim``): the reference's unmodified model classes build their graphs
with these Ops (``tests/golden/make_graph_fixture.py``), and ``tests/graph_worker.py``
differentiates and evaluates them on the GPU against the reference's own values.  All
numerical logic lives in the tested ``Engine``; each Op is a thin adapter.

Ops
---
``SpinTempOp``            physics.calc_spin_temp (physics.py:58-106)
``PseudoVoigtOp``         physics.calc_pseudo_voigt (physics.py:265-309)
``RadiativeTransferOp``   physics.radiative_transfer (physics.py:312-365)
``HISpectraOp``           the fused add_likelihood() graph up to the two ``mu`` vectors
                          (emission_absorption_model.py:69-99 and siblings): keeps the
                          reference's ``pm.Normal(observed=...)`` nodes intact
``HILogLikeOp``           the same plus the Normal logp, as one scalar (``pm.Potential``)
"""

import numpy as np

try:
    import pytensor.tensor as pt
    from pytensor.gradient import DisconnectedType
    from pytensor.graph.basic import Apply, Variable
    from pytensor.graph.op import Op

    HAVE_PYTENSOR = True
except Exception:
    HAVE_PYTENSOR = False
    pt = None
    Variable = ()

    class Op:  # minimal stand-in so the classes below import and ``perform`` can be tested
        __props__ = ()

        def __call__(self, *inputs):
            raise RuntimeError(
                "PyTensor is not installed: symbolic use of caribou_hi_b200 Ops is unavailable; "
                "call .perform(None, inputs, output_storage) or use caribou_hi_b200.Engine directly"
            )

    def Apply(*a, **k):
        raise RuntimeError("PyTensor is not installed")


PHYSICS_INPUTS = ("velocity", "fwhm2", "fwhm_L", "tau_total", "tspin", "filling_factor", "absorption_weight")


def is_variable(x):
    return HAVE_PYTENSOR and isinstance(x, Variable)


def _vec(x, n=None):
    a = np.asarray(x, dtype=np.float64)
    if n is not None:
        a = np.broadcast_to(a, (n,))
    return np.ascontiguousarray(a)


def _physics_engine():
    from . import physics

    return physics._default_engine()


# ---- symbolic helpers for the trivially elementwise functions (pragma: PyTensor only) ----
def gaussian_graph(x, center, fwhm):
    return pt.exp(-4.0 * np.log(2.0) * (x - center) ** 2.0 / fwhm**2.0) * pt.sqrt(
        4.0 * np.log(2.0) / (np.pi * fwhm**2.0)
    )


def log10_column_density_graph(tau_total, spin_temp):
    return pt.log10(1.82243e18 * spin_temp * tau_total)


def log10_nonthermal_pressure_graph(log10_density, fwhm2_nonthermal):
    return np.log10(671.8) + log10_density + pt.log10(fwhm2_nonthermal)


# ---- spin temperature ------------------------------------------------------------------
class SpinTempOp(Op):
    """``tspin = calc_spin_temp(kinetic_temp, density, n_alpha)`` elementwise."""

    __props__ = ()

    def make_node(self, kinetic_temp, density, n_alpha):
        ins = [pt.as_tensor_variable(x) for x in (kinetic_temp, density, n_alpha)]
        ins = list(pt.broadcast_arrays(*ins))
        return Apply(self, ins, [ins[0].type()])

    def perform(self, node, inputs, output_storage):
        tk, de, na = inputs
        output_storage[0][0] = _physics_engine().spin_temp(tk, de, na)

    def infer_shape(self, fgraph, node, input_shapes):
        return [input_shapes[0]]

    def L_op(self, inputs, outputs, output_grads):
        return SpinTempGradOp()(*inputs, output_grads[0])


class SpinTempGradOp(Op):
    __props__ = ()

    def make_node(self, kinetic_temp, density, n_alpha, gbar):
        ins = [pt.as_tensor_variable(x) for x in (kinetic_temp, density, n_alpha, gbar)]
        return Apply(self, ins, [ins[0].type(), ins[1].type(), ins[2].type()])

    def perform(self, node, inputs, output_storage):
        tk, de, na, gbar = inputs
        g = _physics_engine().spin_temp_vjp(tk, de, na, gbar)
        for i in range(3):
            output_storage[i][0] = g[i]


# ---- pseudo-Voigt profile -----------------------------------------------------------------
class PseudoVoigtOp(Op):
    """``profile[S, N] = calc_pseudo_voigt(velo_axis, velocity, fwhm2, fwhm_L)``; the
    velocity axis is data (a constant of the Op)."""

    __props__ = ("axis_key",)

    def __init__(self, velo_axis):
        self.velo_axis = np.ascontiguousarray(velo_axis, dtype=np.float64)
        self.axis_key = self.velo_axis.tobytes()

    def make_node(self, velocity, fwhm2, fwhm_L):
        velocity = pt.as_tensor_variable(velocity)
        fwhm2 = pt.as_tensor_variable(fwhm2)
        fwhm_L = pt.as_tensor_variable(fwhm_L) + pt.zeros_like(fwhm2)  # scalar or per cloud
        return Apply(self, [velocity, fwhm2, fwhm_L], [pt.dmatrix()])

    def perform(self, node, inputs, output_storage):
        velocity, fwhm2, fwhm_L = inputs
        n = np.asarray(velocity).shape[0]
        output_storage[0][0] = _physics_engine().pseudo_voigt(self.velo_axis, _vec(velocity), _vec(fwhm2, n),
                                                              _vec(fwhm_L, n))

    def infer_shape(self, fgraph, node, input_shapes):
        return [(self.velo_axis.shape[0], input_shapes[0][0])]

    def L_op(self, inputs, outputs, output_grads):
        return PseudoVoigtGradOp(self.velo_axis)(*inputs, output_grads[0])


class PseudoVoigtGradOp(Op):
    __props__ = ("axis_key",)

    def __init__(self, velo_axis):
        self.velo_axis = np.ascontiguousarray(velo_axis, dtype=np.float64)
        self.axis_key = self.velo_axis.tobytes()

    def make_node(self, velocity, fwhm2, fwhm_L, gbar):
        ins = [pt.as_tensor_variable(x) for x in (velocity, fwhm2, fwhm_L, gbar)]
        return Apply(self, ins, [ins[0].type(), ins[1].type(), ins[2].type()])

    def perform(self, node, inputs, output_storage):
        velocity, fwhm2, fwhm_L, gbar = inputs
        n = np.asarray(velocity).shape[0]
        g = _physics_engine().pseudo_voigt_vjp(self.velo_axis, _vec(velocity), _vec(fwhm2, n), _vec(fwhm_L, n), gbar)
        for i in range(3):
            output_storage[i][0] = g[i]


# ---- radiative transfer --------------------------------------------------------------------
class RadiativeTransferOp(Op):
    """``tb[S] = radiative_transfer(tau[S, N], tspin[N], filling_factor, bg_temp)``."""

    __props__ = ("bg_temp",)

    def __init__(self, bg_temp):
        self.bg_temp = float(bg_temp)

    def make_node(self, tau, tspin, filling_factor):
        tau = pt.as_tensor_variable(tau)
        tspin = pt.as_tensor_variable(tspin)
        filling_factor = pt.as_tensor_variable(filling_factor) + pt.zeros_like(tspin)
        return Apply(self, [tau, tspin, filling_factor], [pt.dvector()])

    def perform(self, node, inputs, output_storage):
        tau, tspin, ff = inputs
        n = np.asarray(tau).shape[1]
        output_storage[0][0] = _physics_engine().radiative_transfer(tau, _vec(tspin, n), _vec(ff, n), self.bg_temp)

    def infer_shape(self, fgraph, node, input_shapes):
        return [(input_shapes[0][0],)]

    def L_op(self, inputs, outputs, output_grads):
        return RadiativeTransferGradOp(self.bg_temp)(*inputs, output_grads[0])


class RadiativeTransferGradOp(Op):
    __props__ = ("bg_temp",)

    def __init__(self, bg_temp):
        self.bg_temp = float(bg_temp)

    def make_node(self, tau, tspin, filling_factor, gbar):
        ins = [pt.as_tensor_variable(x) for x in (tau, tspin, filling_factor, gbar)]
        return Apply(self, ins, [ins[0].type(), ins[1].type(), ins[2].type()])

    def perform(self, node, inputs, output_storage):
        tau, tspin, ff, gbar = inputs
        n = np.asarray(tau).shape[1]
        g = _physics_engine().radiative_transfer_vjp(tau, _vec(tspin, n), _vec(ff, n), self.bg_temp, gbar)
        for i in range(3):
            output_storage[i][0] = g[i]


# ---- fused spectra / log-likelihood -----------------------------------------------------------
def physics_engine_for(data, n_clouds, bg_temp=3.77, lorentzian=True, device=0, with_baseline=False,
                       baseline_degree=0):
    """Build the ``MODEL_PHYSICS`` engine an Op needs from a bayes_spec-style ``data`` dict
    (``{"emission": SpecData, "absorption": SpecData}``, either may be missing).  Without
    ``with_baseline`` the predicted spectra exclude the baseline (bayes_spec adds its own
    symbolic ``predict_baseline()``, as the reference does)."""
    from .engine import Engine

    eng = Engine("physics", n_clouds, device=device, lorentzian=lorentzian, bg_temp=bg_temp,
                 has_emission="emission" in data, has_absorption="absorption" in data)
    sets = {}
    for key in ("emission", "absorption"):
        if key in data:
            d = data[key]
            if with_baseline:
                sets[key] = d.as_dataset(baseline_degree)
            else:
                sets[key] = dict(spectral=d.spectral, brightness=d.brightness, noise=d.noise)
    eng.set_spectra(emission=sets.get("emission"), absorption=sets.get("absorption"))
    return eng


def _theta_from_inputs(engine, inputs):
    n = engine.n_clouds
    return engine.pack({name: _vec(x, n)[None, :] for name, x in zip(PHYSICS_INPUTS, inputs)}, B=1)


class HISpectraOp(Op):
    """``(mu_emission[S_e], mu_absorption[S_a]) = f(velocity, fwhm2, fwhm_L, tau_total, tspin,
    filling_factor, absorption_weight)``: line profile -> optical depth -> radiative transfer
    / absorption, fused in one kernel launch per dataset.  Outputs for a dataset the engine
    does not have are empty vectors."""

    __props__ = ("engine_key",)

    def __init__(self, engine):
        self.engine = engine
        self.engine_key = id(engine)

    def make_node(self, *inputs):
        if len(inputs) != len(PHYSICS_INPUTS):
            raise TypeError(f"HISpectraOp takes {PHYSICS_INPUTS}")
        ref = pt.as_tensor_variable(inputs[0])
        ins = [pt.as_tensor_variable(x) + pt.zeros_like(ref) for x in inputs]
        return Apply(self, ins, [pt.dvector(), pt.dvector()])

    def perform(self, node, inputs, output_storage):
        theta = _theta_from_inputs(self.engine, inputs)
        mu_e, mu_a = self.engine.spectra(theta)
        output_storage[0][0] = mu_e[0] if mu_e is not None else np.zeros(0)
        output_storage[1][0] = mu_a[0] if mu_a is not None else np.zeros(0)

    def infer_shape(self, fgraph, node, input_shapes):
        return [(self.engine.S_e,), (self.engine.S_a,)]

    def L_op(self, inputs, outputs, output_grads):
        ge, ga = output_grads
        if isinstance(ge.type, DisconnectedType):
            ge = pt.zeros((self.engine.S_e,))
        if isinstance(ga.type, DisconnectedType):
            ga = pt.zeros((self.engine.S_a,))
        return HISpectraGradOp(self.engine)(*inputs, ge, ga)


class HISpectraGradOp(Op):
    __props__ = ("engine_key",)

    def __init__(self, engine):
        self.engine = engine
        self.engine_key = id(engine)

    def make_node(self, *inputs):
        ins = [pt.as_tensor_variable(x) for x in inputs]
        return Apply(self, ins, [ins[i].type() for i in range(len(PHYSICS_INPUTS))])

    def perform(self, node, inputs, output_storage):
        eng = self.engine
        theta = _theta_from_inputs(eng, inputs[:7])
        ge = np.asarray(inputs[7], dtype=np.float64)
        ga = np.asarray(inputs[8], dtype=np.float64)
        grad, _, _ = eng.spectra_vjp(theta, ge[None, :] if eng.has_emission and ge.size else None,
                                     ga[None, :] if eng.has_absorption and ga.size else None)
        for i in range(len(PHYSICS_INPUTS)):
            output_storage[i][0] = np.array(grad[0, i, :])


class HILogLikeOp(Op):
    """Scalar Normal log-likelihood of both datasets as one Op (inputs: the seven cloud
    vectors followed by the baseline coefficient vectors of the datasets the engine has,
    emission first); ``L_op`` is the fused gradient kernel."""

    __props__ = ("engine_key",)

    def __init__(self, engine):
        self.engine = engine
        self.engine_key = id(engine)

    def _split(self, inputs):
        eng = self.engine
        theta = _theta_from_inputs(eng, inputs[:7])
        rest = list(inputs[7:])
        ce = np.asarray(rest.pop(0), dtype=np.float64)[None, :] if eng.nb_e else None
        ca = np.asarray(rest.pop(0), dtype=np.float64)[None, :] if eng.nb_a else None
        return theta, ce, ca

    def make_node(self, *inputs):
        ins = [pt.as_tensor_variable(x) for x in inputs]
        return Apply(self, ins, [pt.dscalar()])

    def perform(self, node, inputs, output_storage):
        theta, ce, ca = self._split(inputs)
        logp = self.engine.logp_grad(theta, ce, ca, want_grad=False)[0]
        output_storage[0][0] = np.asarray(logp[0])

    def infer_shape(self, fgraph, node, input_shapes):
        return [()]

    def L_op(self, inputs, outputs, output_grads):
        grads = HILogLikeGradOp(self.engine)(*inputs)
        return [output_grads[0] * g for g in grads]


class HILogLikeGradOp(Op):
    __props__ = ("engine_key",)

    def __init__(self, engine):
        self.engine = engine
        self.engine_key = id(engine)

    def make_node(self, *inputs):
        ins = [pt.as_tensor_variable(x) for x in inputs]
        return Apply(self, ins, [x.type() for x in ins])

    def perform(self, node, inputs, output_storage):
        eng = self.engine
        theta, ce, ca = HILogLikeOp._split(self, inputs)
        _, grad, gce, gca = eng.logp_grad(theta, ce, ca)
        for i in range(7):
            output_storage[i][0] = np.array(grad[0, i, :])
        k = 7
        if eng.nb_e:
            output_storage[k][0] = np.array(gce[0])
            k += 1
        if eng.nb_a:
            output_storage[k][0] = np.array(gca[0])
