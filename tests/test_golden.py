"""Committed golden vectors under tests/golden/:
  * `ref_*.npz` + `ref_free_rvs.json` (tests/golden/make_golden_ref.py): outputs of THE REFERENCE'S
    OWN SOURCE (`caribou_hi/*.py`, unmodified) executed in the build container on eager-torch
    stand-ins for PyTensor / PyMC / bayes_spec -- these pin the oracle and the CUDA path to the
    expressions the reference's authors wrote (physics, parameter maps, likelihood, gradients);
  * the other `*.npz` (tests/golden/make_golden.py): the oracle's own outputs on further cases
    (guards the checker against drift).
CPU tests: the oracle reproduces both sets.  GPU tests: the CUDA path reproduces both sets
through the C ABI."""
import json

import glob
import os

import numpy as np
import pytest

from helpers import assert_close, assert_grad_close, engine_for
from oracle import models_ref as mr

HERE = os.path.dirname(os.path.abspath(__file__))
ALL = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
REF_FILES = [f for f in ALL if os.path.basename(f).startswith("ref_")]
FILES = [f for f in ALL if not os.path.basename(f).startswith("ref")]
PHYS = os.path.join(HERE, "golden", "refphys.npz")  # the reference's physics.py, 13 functions


def _load(path):
    z = np.load(path)
    kind = str(z["kind"])
    lor = bool(z["lorentzian"])
    kw = {"prior_fwhm_L": 1.0} if lor else {}
    if kind == "emission_absorption":
        kw["prior_sigma_log10_NHI"] = 0.5
    spec = mr.make_spec(kind, 3, **kw)
    data = {}
    for key in ("emission", "absorption"):
        if f"data_{key}_spectral" in z:
            data[key] = mr.OracleData(z[f"data_{key}_spectral"], z[f"data_{key}_brightness"], z[f"data_{key}_noise"])
    free = {k[5:]: z[k] for k in z.files if k.startswith("free_") and not k.startswith("free_baseline_")}
    base = {k[5:]: z[k] for k in z.files if k.startswith("free_baseline_")}
    return z, spec, data, free, base


def test_fixtures_present():
    assert len(FILES) == 12
    assert len(REF_FILES) == 12


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(path):
    from helpers import oracle_logp_grad

    z, spec, data, free, base = _load(path)
    lp, grads = oracle_logp_grad(spec, data, free, base)
    assert_close(lp, z["logp"], rtol=1e-13, what="logp", axis=None)
    fin = np.isfinite(z["logp"])
    for k, g in grads.items():
        assert_close(g[fin], z[f"grad_{k}"][fin], rtol=1e-12, what=k, axis=None)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_cuda_reproduces_golden(cuda_device, path):
    z, spec, data, free, base = _load(path)
    eng = engine_for(spec, data, baseline_degree=1)
    theta = eng.pack(free)
    lp, grad, gce, gca = eng.logp_grad(theta, base.get("baseline_emission_norm"), base.get("baseline_absorption_norm"))
    assert_close(lp, z["logp"], what="logp", axis=None)
    fin = np.isfinite(z["logp"])
    got = dict(eng.unpack_grad(grad))
    if gce is not None:
        got["baseline_emission_norm"] = gce
    if gca is not None:
        got["baseline_absorption_norm"] = gca
    # per RV vector, 60-digit arbitration of components that miss 1e-10 against the fixture's autograd value
    assert_grad_close(got, {n: z[f"grad_{n}"] for n in got}, fin, case=(spec, data, free, base), what="gradient")
    mu_e, mu_a = eng.spectra(theta, base.get("baseline_emission_norm"), base.get("baseline_absorption_norm"))
    if mu_e is not None:
        assert_close(mu_e, z["mu_emission"], what="mu_e")
    if mu_a is not None:
        assert_close(mu_a, z["mu_absorption"], what="mu_a")
    det = eng.deterministics(theta)
    for k in z.files:
        if k.startswith("det_") and k[4:] in det:
            assert_close(det[k[4:]], z[k], rtol=1e-12, what=k)
    eng.close()


# ---------------------------------------------------------------------------------------------
# fixtures produced by the reference's own source
# ---------------------------------------------------------------------------------------------
REF_IDS = [os.path.basename(f)[:-4] for f in REF_FILES]


@pytest.mark.parametrize("path", REF_FILES, ids=REF_IDS)
def test_oracle_matches_reference_source(path):
    """logp, gradients, predicted spectra and EVERY pm.Deterministic of the reference classes."""
    from helpers import oracle_logp_grad

    z, spec, data, free, base = _load(path)
    lp, grads = oracle_logp_grad(spec, data, free, base)
    assert np.array_equal(np.isfinite(lp), np.isfinite(z["logp"]))
    assert_close(lp, z["logp"], rtol=1e-12, what="logp", axis=None)
    fin = np.isfinite(z["logp"])
    names = sorted(grads)
    assert_close(np.concatenate([grads[k][fin] for k in names], axis=1),
                 np.concatenate([z[f"grad_{k}"][fin] for k in names], axis=1), rtol=1e-11, what="gradient")
    mu = mr.predict(spec, data, free, base)
    for key in data:
        assert_close(mu[key], z[f"mu_{key}"], rtol=1e-12, what=f"mu_{key}")
    det = mr.cloud_inputs(spec, free)
    ref_dets = [k[4:] for k in z.files if k.startswith("det_")]
    missing = [k for k in ref_dets if k not in det]
    assert not missing, f"reference deterministics the oracle does not produce: {missing}"
    for k in ref_dets:
        assert_close(np.broadcast_to(det[k], z[f"det_{k}"].shape), z[f"det_{k}"], rtol=1e-12, what=k)


def test_free_rvs_match_reference_source():
    """Names, distribution families (=> PyMC default transform), parameters and cloud/scalar
    shape of the free RVs the reference's add_priors() creates, against the host-side layout."""
    from caribou_hi_b200 import layout

    with open(os.path.join(HERE, "golden", "ref_free_rvs.json")) as f:
        meta = json.load(f)
    assert len(meta) == 12
    suffix = {"Normal": "", "HalfNormal": "_log__", "ChiSquared": "_log__", "Beta": "_logodds__",
              "Uniform": "_interval__", "TruncatedNormal": "_interval__"}
    for case, rvs in meta.items():
        kind, prof = case.rsplit("_", 1)
        lor = prof == "pv"
        ours = layout.slot_names(kind, lor, free_weight=True)
        theirs = {k: v for k, v in rvs.items() if not k.startswith("baseline_")}
        assert set(ours) == set(theirs), (case, sorted(set(ours) ^ set(theirs)))
        scalars = layout.scalar_rvs(kind)
        spec = mr.make_spec(kind, 3, **({"prior_sigma_log10_NHI": 0.5} if kind == "emission_absorption" else {}))
        for name, info in theirs.items():
            assert layout.value_name(name) == name + suffix[info["dist"]], (case, name, info["dist"])
            assert (name in scalars) == (info["dims"] is None), (case, name)
            if info["dist"] == "Beta":
                key = "prior_tkin_factor" if name.startswith("tkin_factor") else "prior_fwhm2_thermal_fraction"
                assert [info["params"]["alpha"], info["params"]["beta"]] == list(spec[key]), (case, name)
            if info["dist"] == "TruncatedNormal":
                assert [info["params"]["mu"], info["params"]["sigma"]] == list(spec["prior_nth_fwhm_1pc"])
                assert info["params"]["lower"] == 0.0 and info["params"]["upper"] is None
            if info["dist"] == "Uniform":
                assert (info["params"]["lower"], info["params"]["upper"]) == (0.0, 1.0)
            if info["dist"] in ("HalfNormal",):
                assert info["params"]["sigma"] == 1.0
            if info["dist"] == "ChiSquared":
                assert info["params"]["nu"] == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("path", REF_FILES, ids=REF_IDS)
def test_cuda_matches_reference_source(cuda_device, path):
    z, spec, data, free, base = _load(path)
    eng = engine_for(spec, data, baseline_degree=1)
    theta = eng.pack(free)
    lp, grad, gce, gca = eng.logp_grad(theta, base.get("baseline_emission_norm"), base.get("baseline_absorption_norm"))
    assert np.array_equal(np.isfinite(lp), np.isfinite(z["logp"]))
    assert_close(lp, z["logp"], what="logp", axis=None)
    fin = np.isfinite(z["logp"])
    got = dict(eng.unpack_grad(grad))
    if gce is not None:
        got["baseline_emission_norm"] = gce
    if gca is not None:
        got["baseline_absorption_norm"] = gca
    # per RV vector, 60-digit arbitration of components that miss 1e-10 against the fixture's autograd value
    assert_grad_close(got, {n: z[f"grad_{n}"] for n in got}, fin, case=(spec, data, free, base), what="gradient")
    mu_e, mu_a = eng.spectra(theta, base.get("baseline_emission_norm"), base.get("baseline_absorption_norm"))
    if mu_e is not None:
        assert_close(mu_e, z["mu_emission"], what="mu_e")
    if mu_a is not None:
        assert_close(mu_a, z["mu_absorption"], what="mu_a")
    det = eng.deterministics(theta)
    ref_dets = [k[4:] for k in z.files if k.startswith("det_")]
    missing = [k for k in ref_dets if k not in det]
    assert not missing, f"reference deterministics the device does not produce: {missing}"
    for k in ref_dets:
        assert_close(det[k], z[f"det_{k}"], rtol=1e-12, what=k)
    eng.close()


def _value(x):
    return np.asarray(x.eval() if hasattr(x, "eval") else x, dtype=np.float64)


def _check_physics(P, z, vjps):
    """`P`: a module with the reference's physics API (oracle or drop-in)."""
    x, c, w = z["g_x"], z["g_c"], z["g_w"]
    assert_close(_value(P.gaussian(x, c, w)), z["gaussian"], rtol=1e-12, what="gaussian", axis=None)
    assert_close(_value(P.lorentzian(x, c, w)), z["lorentzian"], rtol=1e-12, what="lorentzian", axis=None)
    ts = _value(P.calc_spin_temp(z["st_tk"], z["st_dens"], z["st_na"]))
    np.testing.assert_allclose(ts, z["st_out"], rtol=1e-12)
    a, b, p3 = z["e_a"], z["e_b"], z["e_p"]
    np.testing.assert_allclose(_value(P.calc_thermal_fwhm2(a)), z["thermal_fwhm2"], rtol=1e-13)
    np.testing.assert_allclose(_value(P.calc_kinetic_temp(a)), z["kinetic_temp"], rtol=1e-13)
    np.testing.assert_allclose(_value(P.calc_nonthermal_fwhm(a, b, p3)), z["nonthermal_fwhm"], rtol=1e-12)
    np.testing.assert_allclose(_value(P.calc_depth_nonthermal(b, b[::-1].copy(), p3)), z["depth_nonthermal"], rtol=1e-12)
    np.testing.assert_allclose(_value(P.calc_log10_column_density(b, a)), z["log10_column_density"], rtol=1e-13)
    np.testing.assert_allclose(_value(P.calc_log10_density(a / 100, b)), z["log10_density"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(_value(P.calc_tau_total(a * 1e18, a[::-1].copy())), z["tau_total"], rtol=1e-13)
    np.testing.assert_allclose(_value(P.calc_log10_nonthermal_pressure(b, a)), z["log10_nonthermal_pressure"],
                               rtol=1e-12, atol=1e-13)
    for tag in ("g", "pv"):
        prof = _value(P.calc_pseudo_voigt(z[f"pv_{tag}_vax"], z[f"pv_{tag}_vel"], z[f"pv_{tag}_fwhm2"],
                                          float(z[f"pv_{tag}_fwhm_L"])))
        assert_close(prof, z[f"pv_{tag}_out"], rtol=1e-12, what=f"pseudo_voigt {tag}", axis=None)
    tb = _value(P.radiative_transfer(z["rt_tau"], z["rt_tspin"], z["rt_ff"], float(z["rt_bg"])))
    assert_close(tb, z["rt_out"], rtol=1e-12, what="radiative_transfer", axis=None)
    if vjps is None:
        return
    g = vjps.spin_temp_vjp(z["st_tk"], z["st_dens"], z["st_na"], z["st_gbar"])
    for got, key in zip(g, ("st_g_tk", "st_g_dens", "st_g_na")):
        # element-wise: a few derivatives are differences of nearly equal terms (autograd's own
        # rounding, 1e-9 relative on values 1e-12 of the array's scale); array-wise at 1e-10
        np.testing.assert_allclose(got, z[key], rtol=1e-8, atol=1e-300)
        assert_close(got, z[key], what=key, axis=None)
    for tag in ("g", "pv"):
        gv, gw, gl = vjps.pseudo_voigt_vjp(z[f"pv_{tag}_vax"], z[f"pv_{tag}_vel"], z[f"pv_{tag}_fwhm2"],
                                           float(z[f"pv_{tag}_fwhm_L"]), z[f"pv_{tag}_gbar"])
        assert_close(gv, z[f"pv_{tag}_g_vel"], what="d/dvelocity", axis=None)
        assert_close(gw, z[f"pv_{tag}_g_fwhm2"], what="d/dfwhm2", axis=None)
        # the reference's fwhm_L is one scalar: its gradient is the sum over clouds
        assert_close(np.sum(gl), z[f"pv_{tag}_g_fwhm_L"], what="d/dfwhm_L", axis=None)
    gt, gs, gf = vjps.radiative_transfer_vjp(z["rt_tau"], z["rt_tspin"], z["rt_ff"], float(z["rt_bg"]), z["rt_gbar"])
    assert_close(gt, z["rt_g_tau"], what="d/dtau", axis=None)
    assert_close(gs, z["rt_g_tspin"], what="d/dtspin", axis=None)
    assert_close(gf, z["rt_g_ff"], what="d/dff", axis=None)


def test_oracle_physics_matches_reference_source():
    from oracle import physics_ref

    _check_physics(physics_ref, np.load(PHYS), None)


@pytest.mark.gpu
def test_cuda_physics_matches_reference_source(cuda_device):
    """The drop-in `physics` module (values) and the C-ABI VJPs against the reference's own
    physics.py + torch autograd."""
    from caribou_hi_b200 import Engine, physics

    eng = Engine("physics", 4)
    _check_physics(physics, np.load(PHYS), eng)
    eng.close()
