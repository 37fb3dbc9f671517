"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN MODEL CLASSES
(`/root/reference/caribou_hi`, imported unmodified from where it lies) on seeded inputs.

PyTensor / PyMC / bayes_spec are not installable in this image, so the reference's source runs
on the stand-ins in tests/golden/ref_shim/ (eager torch float64 for `pytensor.tensor`, value
replay for the `pymc` distributions, closed-form Normal likelihood, the restated bayes_spec
baseline).  What these vectors pin is therefore every expression the reference's authors wrote
-- `physics.py`, all six classes' `add_priors` / `add_deterministics` / `add_likelihood` -- and
their torch-autograd gradients; what they cannot pin is third-party behaviour (PyMC's transforms
and prior densities, bayes_spec's baseline normalisation).  Run in the build container only:

    python tests/golden/make_golden_ref.py
"""
import inspect
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "ref_shim"))  # pytensor / pymc / bayes_spec stand-ins
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pymc as pm  # noqa: E402  (the stand-in)
import caribou_hi as ref  # noqa: E402  (the reference, unmodified)
from helpers import make_case  # noqa: E402
from oracle import models_ref as mr  # noqa: E402

CLASSES = {
    "emission": ref.EmissionModel,
    "absorption": ref.AbsorptionModel,
    "emission_absorption": ref.EmissionAbsorptionModel,
    "emission_physical": ref.EmissionPhysicalModel,
    "absorption_physical": ref.AbsorptionPhysicalModel,
    "emission_absorption_physical": ref.EmissionAbsorptionPhysicalModel,
}
B, N = 6, 3


class Data:
    """SpecData stand-in handed to the reference: the attributes its classes read; the baseline
    is the oracle's restatement of bayes_spec's."""

    def __init__(self, d):
        # torch tensors: `ndarray - Tensor` does not dispatch to torch the way PyTensor's
        # variables take over NumPy operands
        self.spectral = torch.as_tensor(d.spectral)
        self.brightness = torch.as_tensor(d.brightness)
        self.noise = torch.as_tensor(d.noise)
        self._d = d

    def predict_baseline(self, coeffs):
        return self._d.predict_baseline(coeffs)


def accepted_kwargs(cls, method):
    names = set()
    for k in cls.__mro__:
        f = k.__dict__.get(method)
        if f is not None:
            names |= set(inspect.signature(f).parameters)
    return names - {"self", "args", "kwargs"}


def run_reference(kind, spec, data, free, baseline, b):
    cls = CLASSES[kind]
    ctor = {k: spec[k] for k in ("bg_temp", "depth_nth_fwhm_power") if k in accepted_kwargs(cls, "__init__")}
    model = cls({k: Data(d) for k, d in data.items()}, N, baseline_degree=1, **ctor)
    leaves = {name: torch.tensor(v[b], dtype=torch.float64, requires_grad=True)
              for name, v in list(free.items()) + list(baseline.items())}
    pm.VALUES.clear()
    pm.VALUES.update(leaves)
    prior_kwargs = {k: spec[k] for k in accepted_kwargs(cls, "add_priors")
                    if k in spec and k != "prior_baseline_coeffs"}
    model.add_priors(**prior_kwargs)
    model.add_likelihood()
    return model, leaves


def plain(v):
    if v is None:
        return None
    if hasattr(v, "detach"):
        v = v.detach().numpy()
    return np.asarray(v, dtype=float).tolist()


def physics_fixture():
    """All 13 functions of the reference's physics.py on seeded inputs; cotangent-weighted
    gradients (VJPs) of spin temperature, pseudo-Voigt profile and radiative transfer."""
    from caribou_hi import physics as P

    rng = np.random.default_rng(4242)
    T = lambda x, g=False: torch.tensor(np.asarray(x, dtype=np.float64), requires_grad=g)  # noqa: E731
    out = {}
    n = 64
    x = rng.uniform(-30, 30, n); c = rng.normal(0, 5, n); w = rng.uniform(0.5, 20, n)
    out.update(g_x=x, g_c=c, g_w=w, gaussian=P.gaussian(T(x), T(c), T(w)).numpy(),
               lorentzian=P.lorentzian(T(x), T(c), T(w)).numpy())
    # spin temperature on both sides of the Tk = 300 K switch, incl. the boundary itself
    tk = np.concatenate([10 ** rng.uniform(1, 4, n - 3), [300.0, 299.999999, 300.000001]])
    dens = 10 ** rng.uniform(-2, 3, n); na = 10 ** rng.uniform(-9, -3, n)
    tk_t, d_t, na_t = T(tk, True), T(dens, True), T(na, True)
    ts = P.calc_spin_temp(tk_t, d_t, na_t)
    gbar = rng.standard_normal(n)
    (ts * T(gbar)).sum().backward()
    out.update(st_tk=tk, st_dens=dens, st_na=na, st_out=ts.detach().numpy(), st_gbar=gbar,
               st_g_tk=tk_t.grad.numpy(), st_g_dens=d_t.grad.numpy(), st_g_na=na_t.grad.numpy())
    a = 10 ** rng.uniform(0, 4, n); b = rng.uniform(0.1, 3, n); p3 = rng.uniform(0.2, 0.5, n)
    out.update(e_a=a, e_b=b, e_p=p3,
               thermal_fwhm2=P.calc_thermal_fwhm2(T(a)).numpy(), kinetic_temp=P.calc_kinetic_temp(T(a)).numpy(),
               nonthermal_fwhm=P.calc_nonthermal_fwhm(T(a), T(b), T(p3)).numpy(),
               depth_nonthermal=P.calc_depth_nonthermal(T(b), T(b[::-1].copy()), T(p3)).numpy(),
               log10_column_density=P.calc_log10_column_density(T(b), T(a)).numpy(),
               log10_density=P.calc_log10_density(T(a / 100), T(b)).numpy(),
               tau_total=P.calc_tau_total(T(a * 1e18), T(a[::-1].copy())).numpy(),
               log10_nonthermal_pressure=P.calc_log10_nonthermal_pressure(T(b), T(a)).numpy())
    # pseudo-Voigt profile and radiative transfer, S x N
    S, N_ = 96, 4
    for tag, fl in (("g", 0.0), ("pv", 1.3)):
        vax = np.linspace(-40, 40, S)
        vel, f2, fL = T(rng.normal(0, 8, N_), True), T(rng.uniform(1, 60, N_), True), T(fl, True)
        prof = P.calc_pseudo_voigt(T(vax), vel, f2, fL)
        gb = rng.standard_normal((S, N_))
        (prof * T(gb)).sum().backward()
        out.update({f"pv_{tag}_vax": vax, f"pv_{tag}_vel": vel.detach().numpy(), f"pv_{tag}_fwhm2": f2.detach().numpy(),
                    f"pv_{tag}_fwhm_L": np.float64(fl), f"pv_{tag}_out": prof.detach().numpy(), f"pv_{tag}_gbar": gb,
                    f"pv_{tag}_g_vel": vel.grad.numpy(), f"pv_{tag}_g_fwhm2": f2.grad.numpy(),
                    f"pv_{tag}_g_fwhm_L": fL.grad.numpy()})
    tau = T(10 ** rng.uniform(-3, 1.2, (S, N_)), True)
    tsp, ff = T(10 ** rng.uniform(1, 3.5, N_), True), T(rng.uniform(0.05, 1.0, N_), True)
    tb = P.radiative_transfer(tau, tsp, ff, 3.77)
    gb = rng.standard_normal(S)
    (tb * T(gb)).sum().backward()
    out.update(rt_tau=tau.detach().numpy(), rt_tspin=tsp.detach().numpy(), rt_ff=ff.detach().numpy(), rt_bg=np.float64(3.77),
               rt_out=tb.detach().numpy(), rt_gbar=gb, rt_g_tau=tau.grad.numpy(), rt_g_tspin=tsp.grad.numpy(),
               rt_g_ff=ff.grad.numpy())
    np.savez_compressed(os.path.join(HERE, "refphys.npz"), **out)
    print("refphys.npz", len(out), "arrays", flush=True)


def main():
    physics_fixture()
    meta = {}
    for i, (kind, lor) in enumerate((k, lo) for k in mr.KINDS for lo in (False, True)):
        spec, data, free, baseline = make_case(kind, N, 64, 48, B, seed=700 + i, lorentzian=lor, baseline_degree=1)
        out = {"kind": kind, "lorentzian": lor}
        logp = np.zeros(B)
        grads = {k: np.zeros_like(v) for k, v in list(free.items()) + list(baseline.items())}
        mus, dets = {}, {}
        for b in range(B):
            model, leaves = run_reference(kind, spec, data, free, baseline, b)
            m = model.model
            lp = m.logp
            logp[b] = float(lp)
            if torch.isfinite(lp):
                lp.backward()
                for k, t in leaves.items():
                    if t.grad is not None:
                        grads[k][b] = t.grad.numpy()
            for key, (mu, _, _) in m.observed.items():
                mus.setdefault(key, np.zeros((B,) + tuple(mu.shape)))[b] = mu.detach().numpy()
            for name in m.deterministics:
                v = m.named[name].detach().numpy()
                dets.setdefault(name, np.zeros((B, N)))[b] = np.broadcast_to(v, (N,))
            if b == 0:
                meta[f"{kind}_{'pv' if lor else 'gauss'}"] = {
                    name: {"dist": d, "params": {k: plain(v) for k, v in p.items()}, "dims": dims}
                    for name, (d, p, dims) in m.free.items()}
        out["logp"] = logp
        for key, d in data.items():
            out[f"data_{key}_spectral"] = d.spectral
            out[f"data_{key}_brightness"] = d.brightness
            out[f"data_{key}_noise"] = d.noise
            out[f"mu_{key}"] = mus[key]
        for k, v in list(free.items()) + list(baseline.items()):
            out[f"free_{k}"] = v
            out[f"grad_{k}"] = grads[k]
        for k, v in dets.items():
            out[f"det_{k}"] = v
        name = f"ref_{kind}_{'pv' if lor else 'gauss'}.npz"
        np.savez_compressed(os.path.join(HERE, name), **out)
        print(name, "finite draws:", int(np.isfinite(logp).sum()), "/", B, flush=True)
    with open(os.path.join(HERE, "ref_free_rvs.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
