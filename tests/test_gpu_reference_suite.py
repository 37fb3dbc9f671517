"""The reference's own test-suite re-expressed against the drop-in API
(reference tests/test_physics.py, tests/test_hi_model.py, tests/test_hi_physical_model.py):
same inputs, same assertions, ``caribou_hi`` -> ``caribou_hi_b200``.  Everything that the
reference ``.eval()``s runs on the GPU through the C ABI; Op shims are driven through
``perform`` (PyTensor is not installable here)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- tests/test_physics.py
def test_gaussian(cuda_device):
    from caribou_hi_b200 import physics

    gauss = physics.gaussian(1.0, 0.0, 1.0).eval()
    val = np.exp(-4.0 * np.log(2.0)) * np.sqrt(4.0 * np.log(2.0) / np.pi)
    assert pytest.approx(gauss) == val


def test_lorentzian(cuda_device):
    from caribou_hi_b200 import physics

    lorent = physics.lorentzian(1.0, 0.0, 1.0)
    assert pytest.approx(lorent) == 1.0 / (2.0 * np.pi) / 1.25


def test_calc_spin_temp(cuda_device):
    from caribou_hi_b200 import physics

    spin_temp = physics.calc_spin_temp(1000.0, 1000.0, 1.0e-6).eval()
    assert spin_temp == pytest.approx(1000.0, abs=1.0)
    spin_temp = physics.calc_spin_temp(8000.0, 1.0, 1.0e-6).eval()
    assert spin_temp < 8000.0


def test_scalar_helpers(cuda_device):
    from caribou_hi_b200 import physics

    assert 0.2139**2.0 * 1000.0 == pytest.approx(physics.calc_thermal_fwhm2(1000.0))
    assert 8.0 == pytest.approx(physics.calc_nonthermal_fwhm(2.0, 2.0, 2.0))
    assert 2.0 == pytest.approx(physics.calc_depth_nonthermal(8.0, 2.0, 2.0))
    assert np.log10(1.82243e18) == pytest.approx(physics.calc_log10_column_density(1.0, 1.0).eval())
    assert 0.0 == pytest.approx(physics.calc_log10_density(18.489351, 0.0))
    assert 1.0 == pytest.approx(physics.calc_tau_total(1.82243e18, 1.0))
    assert np.log10(671.8) == pytest.approx(physics.calc_log10_nonthermal_pressure(0.0, 1.0).eval())
    assert 1000.0 == pytest.approx(physics.calc_kinetic_temp(physics.calc_thermal_fwhm2(1000.0)))


def test_calc_psuedo_voight(cuda_device):
    from caribou_hi_b200 import physics

    velo_axis = np.linspace(-100.0, 100.0, 101)
    line_profile = physics.calc_pseudo_voigt(velo_axis, np.array([-5.0, 5.0]), np.array([100.0, 200.0]), 0.0).eval()
    assert line_profile.shape == (101, 2)
    assert not np.any(np.isnan(line_profile))


def test_radiative_transfer(cuda_device):
    from caribou_hi_b200 import physics

    velo_axis = np.linspace(-100.0, 100.0, 101)
    optical_depth = np.array([5.0, 3.0]) * physics.calc_pseudo_voigt(
        velo_axis, np.array([-5.0, 5.0]), np.array([100.0, 200.0]), 0.0
    ).eval()
    emission = physics.radiative_transfer(optical_depth, np.array([1000.0, 5000.0]), 1.0, 2.7).eval()
    assert emission.shape == (101,)
    assert not np.any(np.isnan(emission))


# ------------------------------------------- tests/test_hi_model.py, test_hi_physical_model.py
def _data():
    from caribou_hi_b200 import SpecData

    rng = np.random.default_rng(1234)
    emission = SpecData(np.linspace(-20.0, 20.0, 1000), rng.standard_normal(1000), 1.0)
    absorption = SpecData(np.linspace(-20.0, 20.0, 500), rng.standard_normal(500), 1.0)
    return emission, absorption


def test_abstract_bases_cannot_be_instantiated():
    from caribou_hi_b200 import HIModel, HIPhysicalModel, SpecData

    data = {"emission": SpecData(np.linspace(-20.0, 20.0, 1000), np.zeros(1000), 1.0)}
    with pytest.raises(TypeError):
        HIModel(data, 2, baseline_degree=1)
    with pytest.raises(TypeError):
        HIPhysicalModel(data, 2, baseline_degree=1)


def test_absorption_models(cuda_device):
    from caribou_hi_b200 import AbsorptionModel, AbsorptionPhysicalModel

    _, absorption = _data()
    for cls in (AbsorptionModel, AbsorptionPhysicalModel):
        model = cls({"absorption": absorption}, 2, baseline_degree=1)
        model.add_priors()
        model.add_likelihood()
        assert model._validate()


def test_emission_models(cuda_device):
    from caribou_hi_b200 import EmissionModel, EmissionPhysicalModel

    emission, _ = _data()
    for cls in (EmissionModel, EmissionPhysicalModel):
        model = cls({"emission": emission}, 2, baseline_degree=1)
        model.add_priors(prior_fwhm_L=1.0)
        model.add_likelihood()
        assert model._validate()
        assert "fwhm_L_norm" in model.free_RVs


def test_emission_absorption_models(cuda_device):
    from caribou_hi_b200 import EmissionAbsorptionModel, EmissionAbsorptionPhysicalModel

    emission, absorption = _data()
    data = {"emission": emission, "absorption": absorption}
    model = EmissionAbsorptionModel(data, 2, baseline_degree=1)
    model.add_priors()
    model.add_likelihood()
    assert model._validate()
    assert "log10_wt_ff_tkin" not in model.free_RVs  # absorption_weight is the constant 1
    model = EmissionAbsorptionModel(data, 2, baseline_degree=1)
    model.add_priors(prior_sigma_log10_NHI=0.5)
    model.add_likelihood()
    assert model._validate()
    assert "log10_wt_ff_tkin" in model.free_RVs
    model = EmissionAbsorptionPhysicalModel(data, 2, baseline_degree=1)
    model.add_priors()
    model.add_likelihood()
    assert model._validate()


def test_model_class_logp_matches_oracle(cuda_device):
    """The class-level API (dict of named RVs in, dict of named gradients out) against
    the oracle, including the SpecData baseline normalisation."""
    from caribou_hi_b200 import EmissionAbsorptionPhysicalModel
    from helpers import assert_grad_close, assert_close, oracle_logp_grad
    from oracle import models_ref as mr

    emission, absorption = _data()
    model = EmissionAbsorptionPhysicalModel({"emission": emission, "absorption": absorption}, 3,
                                            baseline_degree=2, bg_temp=2.7, depth_nth_fwhm_power=0.4)
    model.add_priors(prior_fwhm_L=1.5, prior_ff_NHI=2.0e21)
    model.add_likelihood()
    point = model.sample_prior(24, np.random.default_rng(7))
    logp, grads = model.logp_dlogp(point)

    spec = mr.make_spec("emission_absorption_physical", 3, bg_temp=2.7, depth_nth_fwhm_power=0.4,
                        prior_fwhm_L=1.5, prior_ff_NHI=2.0e21)
    data = {"emission": mr.OracleData(emission.spectral, emission.brightness, emission.noise),
            "absorption": mr.OracleData(absorption.spectral, absorption.brightness, absorption.noise)}
    free = {k: v for k, v in point.items() if not k.startswith("baseline_")}
    base = {k: v for k, v in point.items() if k.startswith("baseline_")}
    ref_lp, ref_g = oracle_logp_grad(spec, data, free, base)
    assert_close(logp, ref_lp, what="model logp", axis=None)
    fin = np.isfinite(ref_lp)
    names = sorted(grads)
    assert set(names) == set(ref_g)
    assert_grad_close(grads, ref_g, fin, case=(spec, data, free, base), what=f"model grads {names}")
    pred = model.predict(point)
    ref_mu = mr.predict(spec, data, free, base)
    for key in pred:
        assert_close(pred[key], ref_mu[key], what=f"model predict {key}")


# ---------------------------------------------------------------------------- Op shims
def test_ops_perform_and_vjps(cuda_device):
    import torch
    from caribou_hi_b200 import SpecData, ops
    from helpers import assert_close
    from oracle import models_ref as mr
    from oracle import physics_ref as ph

    rng = np.random.default_rng(5)
    N, S = 3, 120
    v = np.linspace(-40, 40, S)
    vel, f2, fl = rng.normal(0, 5, N), rng.uniform(5, 80, N), rng.uniform(0.1, 2, N)

    out = [[None]]
    ops.PseudoVoigtOp(v).perform(None, [vel, f2, fl], out)
    assert_close(out[0][0], ph.calc_pseudo_voigt(v, vel, f2, fl), what="PseudoVoigtOp", axis=0)
    gbar = rng.standard_normal((S, N))
    t = [torch.tensor(x, requires_grad=True) for x in (vel, f2, fl)]
    (ph.calc_pseudo_voigt(torch.tensor(v), *t) * torch.tensor(gbar)).sum().backward()
    out3 = [[None], [None], [None]]
    ops.PseudoVoigtGradOp(v).perform(None, [vel, f2, fl, gbar], out3)
    for o, r, nm in zip(out3, t, ("velocity", "fwhm2", "fwhm_L")):
        assert_close(o[0], r.grad.numpy(), what=f"PseudoVoigtGradOp {nm}", axis=None)

    tau = rng.uniform(0, 3, (S, N))
    ts, ff = rng.uniform(50, 5000, N), rng.uniform(0.2, 1, N)
    ops.RadiativeTransferOp(3.77).perform(None, [tau, ts, ff], out)
    assert_close(out[0][0], ph.radiative_transfer(tau, ts, ff, 3.77), what="RadiativeTransferOp")
    gs = rng.standard_normal(S)
    t = [torch.tensor(x, requires_grad=True) for x in (tau, ts, ff)]
    (ph.radiative_transfer(*t, 3.77) * torch.tensor(gs)).sum().backward()
    ops.RadiativeTransferGradOp(3.77).perform(None, [tau, ts, ff, gs], out3)
    assert_close(out3[0][0], t[0].grad.numpy(), what="RT grad tau", axis=None)
    assert_close(out3[1][0], t[1].grad.numpy(), what="RT grad tspin", axis=None, rtol=1e-9)
    assert_close(out3[2][0], t[2].grad.numpy(), what="RT grad ff", axis=None, rtol=1e-9)

    tk, de, na = rng.uniform(20, 9000, 50), 10 ** rng.uniform(-2, 3, 50), 10 ** rng.uniform(-8, -4, 50)
    ops.SpinTempOp().perform(None, [tk, de, na], out)
    assert_close(out[0][0], ph.calc_spin_temp(tk, de, na), what="SpinTempOp", rtol=1e-13)

    # fused spectra Op + its gradient Op, and the log-likelihood Op
    data = {"emission": SpecData(v, rng.normal(5, 2, S), 0.3), "absorption": SpecData(v[::2], rng.normal(0.2, 0.1, S // 2), 0.02)}
    inputs = [vel, f2, fl, rng.uniform(0.1, 5, N), ts, ff, rng.uniform(0.5, 1.2, N)]
    eng = ops.physics_engine_for(data, N, bg_temp=3.77)
    spec = mr.make_spec("emission_absorption", N)
    odata = {k: mr.OracleData(d.spectral, d.brightness, d.noise) for k, d in data.items()}
    for d in odata.values():
        d._brightness_offset = 0.0  # the Op returns spectra without the baseline
    det = {k: torch.tensor(x, requires_grad=True) for k, x in zip(ops.PHYSICS_INPUTS, inputs)}
    mu = mr.predict_from_inputs(spec, odata, det)
    out2 = [[None], [None]]
    ops.HISpectraOp(eng).perform(None, inputs, out2)
    assert_close(out2[0][0], mu["emission"].detach().numpy(), what="HISpectraOp e")
    assert_close(out2[1][0], mu["absorption"].detach().numpy(), what="HISpectraOp a")
    ge, ga = rng.standard_normal(S), rng.standard_normal(S // 2)
    ((mu["emission"] * torch.tensor(ge)).sum() + (mu["absorption"] * torch.tensor(ga)).sum()).backward()
    out7 = [[None] for _ in range(7)]
    ops.HISpectraGradOp(eng).perform(None, inputs + [ge, ga], out7)
    got = np.stack([o[0] for o in out7])
    ref = np.stack([det[k].grad.numpy() for k in ops.PHYSICS_INPUTS])
    assert_close(got, ref, what="HISpectraGradOp", axis=1)
    eng.close()

    eng = ops.physics_engine_for(data, N, bg_temp=3.77, with_baseline=True, baseline_degree=1)
    odata = {k: mr.OracleData(d.spectral, d.brightness, d.noise) for k, d in data.items()}
    be, ba = rng.standard_normal(2), rng.standard_normal(2)
    det = {k: torch.tensor(x, requires_grad=True) for k, x in zip(ops.PHYSICS_INPUTS, inputs)}
    bt = {"baseline_emission_norm": torch.tensor(be, requires_grad=True),
          "baseline_absorption_norm": torch.tensor(ba, requires_grad=True)}
    lp = mr.logp_from_inputs(spec, odata, det, bt)
    lp.backward()
    out1 = [[None]]
    ops.HILogLikeOp(eng).perform(None, inputs + [be, ba], out1)
    assert_close(out1[0][0], lp.detach().numpy(), what="HILogLikeOp", axis=None)
    out9 = [[None] for _ in range(9)]
    ops.HILogLikeGradOp(eng).perform(None, inputs + [be, ba], out9)
    got = np.concatenate([o[0] for o in out9])
    ref = np.concatenate([det[k].grad.numpy() for k in ops.PHYSICS_INPUTS] + [bt[k].grad.numpy() for k in bt])
    assert_close(got, ref, what="HILogLikeGradOp", axis=None)
    eng.close()
