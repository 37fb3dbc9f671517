"""Traces the PyTensor graph of the reference's UNMODIFIED model classes with the CUDA Ops wired in, and
commits it as tests/golden/graph_*.json (build container only; the GPU box has no /root/reference).

Two wirings, both from INTEGRATION.md:
  functions  `caribou_hi.physics` replaced by `caribou_hi_b200.physics` in every model module (section 2):
             the reference's own add_priors() / add_likelihood() bodies then build SpinTempOp,
             PseudoVoigtOp and RadiativeTransferOp nodes;
  fused      the add_likelihood() patch of section 3 (HISpectraOp keeps the pm.Normal nodes intact).
PyTensor / PyMC / bayes_spec are the symbolic stand-ins of tests/golden/graph_shim/ (README there).
What is committed is the forward graph of the model's summed observed log-density as a function of the free
RVs and baseline coefficients; tests/graph_worker.py rebuilds it on the GPU box, differentiates it through
every Op's L_op and compares logp and gradients with tests/golden/ref_*.npz.

    python tests/golden/make_graph_fixture.py            # (re)write the JSON files
    python tests/golden/make_graph_fixture.py --check    # compare with the committed files
"""
import inspect
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "graph_shim"))  # symbolic pytensor / pymc / bayes_spec
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pymc as pm  # noqa: E402  (stand-in)
import pytensor.tensor as pt  # noqa: E402  (stand-in)
from pytensor.graph.basic import Constant, toposort  # noqa: E402

import caribou_hi as ref  # noqa: E402  (the reference, unmodified)
from caribou_hi_b200 import ops  # noqa: E402
from caribou_hi_b200 import physics as b200_physics  # noqa: E402
from helpers import make_case  # noqa: E402
from oracle import models_ref as mr  # noqa: E402

assert ops.HAVE_PYTENSOR, "the symbolic stand-in must be importable as pytensor"

CLASSES = {
    "emission": ref.EmissionModel,
    "absorption": ref.AbsorptionModel,
    "emission_absorption": ref.EmissionAbsorptionModel,
    "emission_physical": ref.EmissionPhysicalModel,
    "absorption_physical": ref.AbsorptionPhysicalModel,
    "emission_absorption_physical": ref.EmissionAbsorptionPhysicalModel,
}
MODEL_MODULES = ["hi_model", "hi_physical_model", "absorption_model", "emission_model", "emission_absorption_model",
                 "absorption_physical_model", "emission_physical_model", "emission_absorption_physical_model"]
N = 3
# function-level wiring is traced for these (every Op class appears); the fused wiring for all twelve
FUNCTION_CASES = {("emission_absorption_physical", True), ("emission", False), ("absorption", True)}


class Data:
    """What the reference's classes and the bayes_spec stand-in read from a SpecData."""

    def __init__(self, d):
        self.spectral, self.brightness, self.noise = d.spectral, d.brightness, d.noise
        self.spectral_norm = d.spectral_norm
        self.brightness_scale, self.brightness_offset = d._brightness_scale, d._brightness_offset


class StubEngine:
    """Stands in for the Engine while the graph is traced without a GPU: HISpectraOp.make_node / L_op /
    infer_shape only need the channel counts."""

    def __init__(self, data, n_clouds):
        self.n_clouds = n_clouds
        self.has_emission, self.has_absorption = "emission" in data, "absorption" in data
        self.S_e = len(data["emission"].spectral) if self.has_emission else 0
        self.S_a = len(data["absorption"].spectral) if self.has_absorption else 0
        self.nb_e = self.nb_a = 0


def add_likelihood_fused(self):
    """INTEGRATION.md section 3, written once for the six classes (absorption-only models have no tspin /
    filling factor / weight in the likelihood: ones)."""
    engine = ops.physics_engine_for(self.data, self.n_clouds, bg_temp=getattr(self, "bg_temp", 3.77), lorentzian=True)
    m = self.model
    vel = m["velocity"]
    one = pt.ones_like(vel)
    fwhm_L = m["fwhm_L"] + pt.zeros_like(vel)  # scalar -> per cloud
    has_e = "emission" in self.data
    tspin = m["tspin"] if has_e else one
    ff = m["filling_factor"] if "filling_factor" in m else one
    wt = m["absorption_weight"] if "absorption_weight" in m else one
    mu_e, mu_a = ops.HISpectraOp(engine)(vel, m["fwhm2"], fwhm_L, m["tau_total"], tspin, ff, wt)
    baseline = self.predict_baseline()
    with self.model:
        if "absorption" in self.data:
            pm.Normal("absorption", mu=mu_a + baseline["absorption"], sigma=self.data["absorption"].noise,
                      observed=self.data["absorption"].brightness)
        if has_e:
            pm.Normal("emission", mu=mu_e + baseline["emission"], sigma=self.data["emission"].noise,
                      observed=self.data["emission"].brightness)


def accepted_kwargs(cls, method):
    names = set()
    for k in cls.__mro__:
        f = k.__dict__.get(method)
        if f is not None:
            names |= set(inspect.signature(f).parameters)
    return names - {"self", "args", "kwargs"}


def build(kind, lorentzian, mode, seed):
    spec, data, free, baseline = make_case(kind, N, 64, 48, 6, seed=seed, lorentzian=lorentzian, baseline_degree=1)
    cls = CLASSES[kind]
    wrapped = {k: Data(d) for k, d in data.items()}
    # wiring
    saved = {name: getattr(ref, name).physics for name in MODEL_MODULES if hasattr(getattr(ref, name), "physics")}
    saved_al = cls.__dict__.get("add_likelihood")
    saved_factory = ops.physics_engine_for
    try:
        if mode == "functions":
            for name in saved:
                getattr(ref, name).physics = b200_physics
        else:
            cls.add_likelihood = add_likelihood_fused
            ops.physics_engine_for = lambda d, n, **kw: StubEngine(d, n)
        ctor = {k: spec[k] for k in ("bg_temp", "depth_nth_fwhm_power") if k in accepted_kwargs(cls, "__init__")}
        model = cls(wrapped, N, baseline_degree=1, **ctor)
        prior_kwargs = {k: spec[k] for k in accepted_kwargs(cls, "add_priors") if k in spec and k != "prior_baseline_coeffs"}
        model.add_priors(**prior_kwargs)
        model.add_likelihood()
    finally:
        for name, mod in saved.items():
            getattr(ref, name).physics = mod
        ops.physics_engine_for = saved_factory
        if saved_al is not None:
            cls.add_likelihood = saved_al
        elif "add_likelihood" in cls.__dict__:
            del cls.add_likelihood
    return model, wrapped, spec


def describe_op(op, wrapped):
    name = type(op).__name__
    if name == "Elemwise":
        return {"op": "Elemwise", "name": op.name}
    if name == "Sum":
        return {"op": "Sum", "axis": op.axis}
    if name == "Subtensor":
        return {"op": "Subtensor", "index": json.loads(json.dumps(op.frozen))}
    if name == "SpinTempOp":
        return {"op": "SpinTempOp"}
    if name == "PseudoVoigtOp":
        for key, d in wrapped.items():
            if np.array_equal(op.velo_axis, d.spectral):
                return {"op": "PseudoVoigtOp", "axis_of": key}
        raise ValueError("PseudoVoigtOp axis is not a dataset axis")
    if name == "RadiativeTransferOp":
        return {"op": "RadiativeTransferOp", "bg_temp": op.bg_temp}
    if name == "HISpectraOp":
        return {"op": "HISpectraOp", "n_clouds": op.engine.n_clouds, "datasets": sorted(wrapped)}
    raise ValueError(f"unexpected Op {name} in the traced graph")


def serialise(model, wrapped):
    m = model.model
    free = dict(m.free)
    by_id = {id(v): ["in", name] for name, v in free.items()}
    consts = []
    nodes = []

    def ref_of(v):
        if id(v) in by_id:
            return by_id[id(v)]
        if isinstance(v, Constant):
            for key, d in wrapped.items():
                for field in ("spectral", "brightness", "noise", "spectral_norm"):
                    a = getattr(d, field)
                    if v.data.shape == np.shape(a) and v.data.ndim > 0 and np.array_equal(v.data, a):
                        return ["data", key, field]
            consts.append(np.asarray(v.data, dtype=np.float64).tolist() if v.data.dtype != bool else v.data.tolist())
            return ["const", len(consts) - 1]
        raise ValueError(f"variable {v!r} is neither an input, a constant nor a node output")

    for node in toposort([m.logp]):
        ins = [ref_of(i) for i in node.inputs]
        d = describe_op(node.op, wrapped)
        d["in"] = ins
        d["n_out"] = len(node.outputs)
        nodes.append(d)
        for k, o in enumerate(node.outputs):
            by_id[id(o)] = ["node", len(nodes) - 1, k]
    return {"inputs": {name: {"ndim": v.ndim, "dist": m.free_info[name][0], "dims": m.free_info[name][1]} for name, v in free.items()},
            "deterministics": {name: by_id[id(m.named[name])] for name in m.deterministics if id(m.named[name]) in by_id},
            "observed_mu": {name: by_id[id(mu)] for name, mu in m.observed.items()},
            "consts": consts, "nodes": nodes, "output": by_id[id(m.logp)]}


def main():
    check = "--check" in sys.argv
    bad = 0
    for i, (kind, lor) in enumerate((k, lo) for k in mr.KINDS for lo in (False, True)):
        for mode in ("fused", "functions"):
            if mode == "functions" and (kind, lor) not in FUNCTION_CASES:
                continue
            model, wrapped, spec = build(kind, lor, mode, 700 + i)
            g = serialise(model, wrapped)
            g["kind"], g["lorentzian"], g["mode"], g["bg_temp"] = kind, lor, mode, float(spec["bg_temp"])
            g["reference_class"] = CLASSES[kind].__name__
            name = f"graph_{kind}_{'pv' if lor else 'gauss'}_{mode}.json"
            text = json.dumps(g, sort_keys=True)
            path = os.path.join(HERE, name)
            ops_used = sorted({n["op"] + (":" + n["name"] if n["op"] == "Elemwise" else "") for n in g["nodes"]})
            if check:
                same = os.path.exists(path) and open(path).read() == text
                bad += not same
                print(name, "same" if same else "DIFFERS", flush=True)
            else:
                open(path, "w").write(text)
                print(name, len(g["nodes"]), "nodes;", ", ".join(o for o in ops_used if not o.startswith("Elemwise")), flush=True)
    if check and bad:
        raise SystemExit(f"{bad} traced graphs differ from the committed fixtures")


if __name__ == "__main__":
    main()
