"""GPU half of the PyTensor-Op tests (run as a subprocess by tests/test_gpu_graph.py so that the symbolic
stand-in of tests/golden/graph_shim/ can be imported as `pytensor` without touching the other tests).

For every committed tests/golden/graph_*.json -- the graph the reference's UNMODIFIED model class builds with
the CUDA Ops wired in (tests/golden/make_graph_fixture.py) -- rebuild it with the real Ops of
caribou_hi_b200.ops, differentiate it through every Op's L_op, evaluate logp and gradients through perform on
the GPU, and compare with what the reference's own expressions give (tests/golden/ref_*.npz) at 1e-10."""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden", "graph_shim"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pytensor  # noqa: E402  (the stand-in)
import pytensor.tensor as pt  # noqa: E402
from pytensor.gradient import grad  # noqa: E402
from pytensor.graph.basic import Constant, toposort  # noqa: E402

from caribou_hi_b200 import ops  # noqa: E402
from helpers import assert_close, assert_grad_close, make_case  # noqa: E402
from oracle import models_ref as mr  # noqa: E402

assert ops.HAVE_PYTENSOR
GOLD = os.path.join(ROOT, "tests", "golden")
OUR_OPS = (ops.SpinTempOp, ops.PseudoVoigtOp, ops.RadiativeTransferOp, ops.HISpectraOp, ops.HILogLikeOp,
           ops.SpinTempGradOp, ops.PseudoVoigtGradOp, ops.RadiativeTransferGradOp, ops.HISpectraGradOp, ops.HILogLikeGradOp)


class Data:
    def __init__(self, spectral, brightness, noise):
        self.spectral, self.brightness, self.noise = spectral, brightness, noise


def rebuild(g, data):
    """The traced graph with the real Ops; returns (input variables, logp variable, observed mu variables)."""
    inputs = {name: pt.tensor(name, ndim=info["ndim"]) for name, info in g["inputs"].items()}
    outs = []
    engines = []

    def var(r):
        if r[0] == "in":
            return inputs[r[1]]
        if r[0] == "const":
            return Constant(np.asarray(g["consts"][r[1]]))
        if r[0] == "data":
            d = data[r[1]]
            if r[2] == "spectral_norm":
                return Constant((d.spectral - d.spectral.mean()) / d.spectral.std())
            return Constant(getattr(d, r[2]))
        return outs[r[1]][r[2]]

    for n in g["nodes"]:
        kind = n["op"]
        if kind == "Elemwise":
            op = pt.Elemwise.REGISTRY[n["name"]]
        elif kind == "Sum":
            op = pt.Sum(n["axis"])
        elif kind == "Subtensor":
            op = pt.Subtensor(pt._thaw(tuple(tuple(x) for x in n["index"])))
        elif kind == "SpinTempOp":
            op = ops.SpinTempOp()
        elif kind == "PseudoVoigtOp":
            op = ops.PseudoVoigtOp(data[n["axis_of"]].spectral)
        elif kind == "RadiativeTransferOp":
            op = ops.RadiativeTransferOp(n["bg_temp"])
        elif kind == "HISpectraOp":
            eng = ops.physics_engine_for({k: data[k] for k in n["datasets"]}, n["n_clouds"], bg_temp=g["bg_temp"], lorentzian=True)
            engines.append(eng)
            op = ops.HISpectraOp(eng)
        else:
            raise ValueError(kind)
        res = op(*[var(r) for r in n["in"]])
        outs.append(res if isinstance(res, list) else [res])
    mus = {k: var(r) for k, r in g["observed_mu"].items()}
    return inputs, var(g["output"]), mus, engines


def check_graph(path):
    g = json.load(open(path))
    tag = f"{g['kind']}_{'pv' if g['lorentzian'] else 'gauss'}"
    z = np.load(os.path.join(GOLD, f"ref_{tag}.npz"), allow_pickle=True)
    data = {k: Data(z[f"data_{k}_spectral"], z[f"data_{k}_brightness"], z[f"data_{k}_noise"])
            for k in ("emission", "absorption") if f"data_{k}_spectral" in z.files}
    inputs, logp, mus, engines = rebuild(g, data)
    names = list(inputs)
    grads = grad(logp, [inputs[n] for n in names])  # runs every Op's L_op
    mu_names = sorted(mus)
    fn = pytensor.function([inputs[n] for n in names], [logp] + grads + [mus[k] for k in mu_names])
    B = z["logp"].shape[0]
    got_lp = np.zeros(B)
    got = {n: np.zeros_like(z[f"grad_{n}"]) for n in names}
    for b in range(B):
        vals = [z[f"free_{n}"][b] if inputs[n].ndim else z[f"free_{n}"][b, 0] for n in names]
        res = fn(*vals)
        got_lp[b] = res[0]
        for n, gv in zip(names, res[1 : 1 + len(names)]):
            got[n][b] = gv
        for k, mu in zip(mu_names, res[1 + len(names) :]):
            assert_close(mu, z[f"mu_{k}"][b], what=f"{os.path.basename(path)} mu_{k}[{b}]")
    assert_close(got_lp, z["logp"], what=f"{os.path.basename(path)} logp", axis=None)
    fin = np.isfinite(z["logp"])
    # the seeded case the fixture was made from (tests/golden/make_golden_ref.py), for the 60-digit arbiter
    i = [(k, lo) for k in mr.KINDS for lo in (False, True)].index((g["kind"], g["lorentzian"]))
    case = make_case(g["kind"], 3, 64, 48, B, seed=700 + i, lorentzian=g["lorentzian"], baseline_degree=1)
    assert_grad_close(got, {n: z[f"grad_{n}"] for n in names}, fin, case=case, what=os.path.basename(path))
    # infer_shape of every CUDA Op against the shapes perform produced
    n_ours = 0
    order = toposort([logp] + grads)
    probe = pytensor.function([inputs[n] for n in names], [o for node in order for o in node.outputs] +
                              [i for node in order for i in node.inputs if not isinstance(i, Constant)])
    vals = [z[f"free_{n}"][0] if inputs[n].ndim else z[f"free_{n}"][0, 0] for n in names]
    flat = probe(*vals)
    n_out = sum(len(node.outputs) for node in order)
    out_vals = iter(flat[:n_out])
    in_vals = iter(flat[n_out:])
    for node in order:
        o_shapes = [np.shape(next(out_vals)) for _ in node.outputs]
        i_shapes = [np.shape(i.data) if isinstance(i, Constant) else np.shape(next(in_vals)) for i in node.inputs]
        if isinstance(node.op, OUR_OPS):
            n_ours += 1
            if type(node.op).__name__.endswith("GradOp"):
                continue
            inferred = node.op.infer_shape(None, node, i_shapes)
            assert [tuple(s) for s in inferred] == [tuple(s) for s in o_shapes], (type(node.op).__name__, inferred, o_shapes)
    for e in engines:
        e.close()
    return n_ours, int(fin.sum())


def check_loglike_op():
    """HILogLikeOp (scalar logp with the Normal folded into the kernel, for pm.Potential): value and
    gradients through make_node / perform / L_op against Engine.logp_grad."""
    z = np.load(os.path.join(GOLD, "ref_emission_absorption_pv.npz"), allow_pickle=True)

    class D(Data):
        def as_dataset(self, degree):
            x = (self.spectral - self.spectral.mean()) / self.spectral.std()
            scale = float(np.median(self.noise))
            return dict(spectral=self.spectral, brightness=self.brightness, noise=self.noise,
                        basis=np.stack([x**i / (i + 1.0) ** i * scale for i in range(degree + 1)]),
                        offset=float(np.median(self.brightness)))

    data = {k: D(z[f"data_{k}_spectral"], z[f"data_{k}_brightness"], z[f"data_{k}_noise"]) for k in ("emission", "absorption")}
    eng = ops.physics_engine_for(data, 3, bg_temp=3.77, lorentzian=True, with_baseline=True, baseline_degree=1)
    names = list(ops.PHYSICS_INPUTS)
    vs = [pt.dvector(n) for n in names] + [pt.dvector("ce"), pt.dvector("ca")]
    op = ops.HILogLikeOp(eng)
    assert op == ops.HILogLikeOp(eng) and hash(op) == hash(ops.HILogLikeOp(eng)) and op != ops.HISpectraOp(eng)
    lp = op(*vs)
    gs = grad(lp * 2.0, vs)  # the factor checks that L_op scales by the output gradient
    fn = pytensor.function(vs, [lp] + gs)
    b = int(np.nonzero(np.isfinite(z["logp"]))[0][0])
    det = {n: z[f"det_{n}"][b] for n in names}
    vals = [det[n] for n in names] + [z["free_baseline_emission_norm"][b], z["free_baseline_absorption_norm"][b]]
    res = fn(*vals)
    theta = eng.pack({n: det[n][None, :] for n in names}, B=1)
    ref_lp, ref_g, gce, gca = eng.logp_grad(theta, vals[-2][None, :], vals[-1][None, :])
    assert abs(res[0] - ref_lp[0]) <= 1e-12 * abs(ref_lp[0])
    assert abs(ref_lp[0] - z["logp"][b]) <= 1e-10 * abs(z["logp"][b])
    ug = eng.unpack_grad(ref_g)
    for n, gv in zip(names, res[1:8]):
        assert_close(gv, 2.0 * ug[n][0], rtol=1e-12, what=f"HILogLikeOp d/d{n}")
    assert_close(res[8], 2.0 * gce[0], rtol=1e-12, what="HILogLikeOp d/dce")
    assert_close(res[9], 2.0 * gca[0], rtol=1e-12, what="HILogLikeOp d/dca")
    assert op.infer_shape(None, None, []) == [()]
    eng.close()


def main():
    paths = sorted(glob.glob(os.path.join(GOLD, "graph_*.json")))
    if len(sys.argv) > 1:
        paths = [p for p in paths if any(a in p for a in sys.argv[1:])]
    total = 0
    for p in paths:
        n_ours, n_fin = check_graph(p)
        total += 1
        print(f"{os.path.basename(p)}: ok ({n_ours} CUDA Op nodes incl. gradient Ops, {n_fin} finite draws)", flush=True)
    check_loglike_op()
    print(f"GRAPH_WORKER_OK {total}")


if __name__ == "__main__":
    main()
