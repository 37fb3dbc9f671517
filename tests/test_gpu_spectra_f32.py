"""The optional single-precision path (predicted spectra, chi_spectra_f32) against the fp64
oracle at the north star's fp32 bar: 1e-5 relative, with the comparison rule of SURVEY.md
section 8c (atol = rtol * max|ref| over each spectrum)."""

import numpy as np
import pytest

from helpers import assert_close, engine_for, make_case
from oracle import models_ref as mr

pytestmark = pytest.mark.gpu

RTOL32 = 1e-5  # BASELINE.json north_star: 1e-5 in the optional fp32 path
KINDS = list(mr.KINDS)


def _check(kind, N, S_e, S_a, B, seed, lorentzian, baseline_degree=1, shared_axis=False, **kw):
    spec, data, free, baseline = make_case(kind, N, S_e, S_a, B, seed, lorentzian, baseline_degree,
                                           shared_axis=shared_axis, **kw)
    eng = engine_for(spec, data, baseline_degree)
    ref = mr.predict(spec, data, free, baseline)
    ce, ca = baseline.get("baseline_emission_norm"), baseline.get("baseline_absorption_norm")
    mu_e, mu_a = eng.spectra(eng.pack(free), ce, ca, dtype=np.float32)
    n_ok = 0
    if mu_e is not None:
        assert mu_e.dtype == np.float32
        assert_close(mu_e, ref["emission"], rtol=RTOL32, what=f"{kind} f32 mu_e")
        n_ok = int(np.isfinite(ref["emission"]).all(axis=1).sum())
    if mu_a is not None:
        assert mu_a.dtype == np.float32
        assert_close(mu_a, ref["absorption"], rtol=RTOL32, what=f"{kind} f32 mu_a")
        n_ok = int(np.isfinite(ref["absorption"]).all(axis=1).sum())
    eng.close()
    return n_ok


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("lorentzian", [False, True])
def test_f32_spectra_all_models(cuda_device, kind, lorentzian):
    assert _check(kind, 3, 200, 100, 64, seed=21, lorentzian=lorentzian) >= 32


@pytest.mark.parametrize("kind", ["emission_absorption", "emission_absorption_physical"])
@pytest.mark.parametrize("lorentzian", [False, True])
@pytest.mark.parametrize("N", [1, 4, 8, 11])
def test_f32_spectra_shared_axis_one_pass(cuda_device, kind, lorentzian, N):
    _check(kind, N, 301, 301, 48, seed=3 + N, lorentzian=lorentzian, shared_axis=True)


def test_f32_spectra_odd_even_channel_counts_and_baselines(cuda_device):
    for S_e, S_a, deg in [(2, 3, 0), (257, 64, 2), (1000, 999, 3)]:
        _check("emission_absorption", 2, S_e, S_a, 16, seed=S_e, lorentzian=True, baseline_degree=deg)
        _check("emission_absorption", 2, S_e, S_e, 16, seed=S_e + 1, lorentzian=False, baseline_degree=deg,
               shared_axis=True)


def _physics_case(inp, ve, va, lorentzian=True, bg=3.77):
    """Likelihood inputs directly (CHI_MODEL_PHYSICS): f32 spectra against the oracle."""
    from caribou_hi_b200 import Engine

    N = inp["velocity"].shape[1]
    data = {"emission": mr.OracleData(ve, np.zeros_like(ve), 0.1), "absorption": mr.OracleData(va, np.zeros_like(va), 0.01)}
    spec = mr.make_spec("emission_absorption", N, bg_temp=bg)
    ref = mr.predict_from_inputs(spec, data, {k: np.asarray(v) for k, v in inp.items()})
    eng = Engine("physics", N, lorentzian=lorentzian, bg_temp=bg)
    eng.set_spectra(emission=dict(spectral=ve, brightness=np.zeros_like(ve), noise=0.1),
                    absorption=dict(spectral=va, brightness=np.zeros_like(va), noise=0.01))
    theta = eng.pack(inp)
    mu_e, mu_a = eng.spectra(theta, dtype=np.float32)
    mu_e64, mu_a64 = eng.spectra(theta)
    eng.close()
    ref_e = np.asarray(ref["emission"]) - data["emission"]._brightness_offset
    ref_a = np.asarray(ref["absorption"]) - data["absorption"]._brightness_offset
    return (mu_e, mu_a), (mu_e64, mu_a64), (ref_e, ref_a)


@pytest.mark.parametrize("shared", [False, True])
def test_f32_warm_optically_thin_gas(cuda_device, shared):
    """tau ~ 1e-4..1e-2 with Ts ~ 1e3..1e4 K: 1 - exp(-tau) must not be formed by cancellation."""
    rng = np.random.default_rng(5)
    B, N = 64, 4
    ve = np.linspace(-100, 100, 1024)
    va = ve if shared else np.linspace(-40, 40, 500)
    inp = {
        "velocity": rng.normal(0, 20, (B, N)),
        "fwhm2": rng.uniform(100, 900, (B, N)),
        "fwhm_L": rng.uniform(0.0, 2.0, (B, N)),
        "tau_total": 10 ** rng.uniform(-3, -1, (B, N)),  # integrated: peak tau = tau_total * phi ~ 1e-4..1e-2
        "tspin": 10 ** rng.uniform(3, 4, (B, N)),
        "filling_factor": rng.uniform(0.2, 1.0, (B, N)),
        "absorption_weight": rng.uniform(0.5, 1.0, (B, N)),
    }
    (e32, a32), (e64, a64), _ = _physics_case(inp, ve, va)
    assert_close(e32, e64, rtol=RTOL32, what="thin warm gas, emission")
    assert_close(a32, a64, rtol=RTOL32, what="thin warm gas, absorption")
    assert np.max(np.abs(a64)) < 0.05  # the case is what it claims to be


@pytest.mark.parametrize("shared", [False, True])
def test_f32_unresolved_lines_far_from_zero_velocity(cuda_device, shared):
    """Lines as narrow as one channel at |v| ~ 100 km/s: v - c must not be rounded to float first."""
    rng = np.random.default_rng(6)
    B, N = 64, 3
    ve = np.linspace(-130, 130, 4096)
    va = ve if shared else np.linspace(-128, 128, 2048)
    inp = {
        "velocity": rng.choice([-1.0, 1.0], (B, N)) * rng.uniform(90, 125, (B, N)),
        "fwhm2": 10 ** rng.uniform(-6, -2, (B, N)),  # W = the channel width
        "fwhm_L": np.zeros((B, N)),
        "tau_total": rng.uniform(0.05, 2.0, (B, N)),
        "tspin": rng.uniform(20, 200, (B, N)),
        "filling_factor": rng.uniform(0.5, 1.0, (B, N)),
        "absorption_weight": np.ones((B, N)),
    }
    (e32, a32), (e64, a64), _ = _physics_case(inp, ve, va, lorentzian=False)
    assert_close(e32, e64, rtol=RTOL32, what="unresolved lines, emission")
    assert_close(a32, a64, rtol=RTOL32, what="unresolved lines, absorption")


def test_f32_opaque_and_mixed_gas_against_the_oracle(cuda_device):
    rng = np.random.default_rng(8)
    B, N = 48, 5
    ve = np.linspace(-60, 60, 777)
    va = np.linspace(-30, 30, 333)
    inp = {
        "velocity": rng.normal(0, 10, (B, N)),
        "fwhm2": 10 ** rng.uniform(0, 3, (B, N)),
        "fwhm_L": rng.uniform(0.0, 5.0, (B, N)),
        "tau_total": 10 ** rng.uniform(-2, 3, (B, N)),
        "tspin": 10 ** rng.uniform(1, 4, (B, N)),
        "filling_factor": rng.uniform(0.0, 1.0, (B, N)),
        "absorption_weight": rng.uniform(0.1, 1.5, (B, N)),
    }
    inp["filling_factor"][:4] = 1.0
    (e32, a32), _, (ref_e, ref_a) = _physics_case(inp, ve, va)
    assert_close(e32, ref_e, rtol=RTOL32, what="mixed gas, emission vs oracle")
    assert_close(a32, ref_a, rtol=RTOL32, what="mixed gas, absorption vs oracle")


def test_f32_negative_optical_depth(cuda_device):
    """Physics-level callers may hand over tau_total < 0 (amplification): both branches of 1 - 2^x."""
    rng = np.random.default_rng(12)
    B, N = 32, 3
    ve = np.linspace(-60, 60, 512)
    va = np.linspace(-30, 30, 256)
    inp = {
        "velocity": rng.normal(0, 10, (B, N)),
        "fwhm2": rng.uniform(10, 300, (B, N)),
        "fwhm_L": rng.uniform(0.0, 2.0, (B, N)),
        "tau_total": -(10 ** rng.uniform(-2, 1.3, (B, N))),
        "tspin": rng.uniform(30, 300, (B, N)),
        "filling_factor": rng.uniform(0.2, 1.0, (B, N)),
        "absorption_weight": rng.uniform(0.5, 1.0, (B, N)),
    }
    (e32, a32), (e64, a64), (ref_e, ref_a) = _physics_case(inp, ve, va)
    assert_close(e64, ref_e, what="negative tau, fp64 emission vs oracle")
    assert_close(e32, e64, rtol=RTOL32, what="negative tau, emission")
    assert_close(a32, a64, rtol=RTOL32, what="negative tau, absorption")


def test_f32_device_pointer_form_and_sub_batches(cuda_device, monkeypatch):
    import torch

    spec, data, free, baseline = make_case("emission_absorption_physical", 4, 500, 500, 300, seed=9,
                                           lorentzian=True, baseline_degree=0, shared_axis=True)
    eng = engine_for(spec, data, 0)
    theta = eng.pack(free)
    ce, ca = baseline["baseline_emission_norm"], baseline["baseline_absorption_norm"]
    h_e, h_a = eng.spectra(theta, ce, ca, dtype=np.float32)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    d_e = torch.empty((300, 500), dtype=torch.float32, device="cuda")
    d_a = torch.empty((300, 500), dtype=torch.float32, device="cuda")
    eng.spectra_f32_dev(300, t(theta), d_e, d_a, t(ce), t(ca))
    eng.synchronize()
    # same kernels, same arithmetic: bit-identical, NaN draws included
    assert np.array_equal(d_e.cpu().numpy(), h_e, equal_nan=True)
    assert np.array_equal(d_a.cpu().numpy(), h_a, equal_nan=True)
    # one output only
    only_e = torch.empty_like(d_e)
    eng.spectra_f32_dev(300, t(theta), only_e, None, t(ce), None)
    eng.synchronize()
    ref = mr.predict(spec, data, free, baseline)
    assert_close(only_e.cpu().numpy(), ref["emission"], rtol=RTOL32, what="emission only")
    eng.close()


def test_f32_config5_shape_against_fp64_kernels(cuda_device):
    """BASELINE configs[4] per-draw shape (8 clouds, 4096 + 4096 channels, pseudo-Voigt): the
    single-precision spectra of 2048 prior draws against the fp64 kernels."""
    import torch
    from caribou_hi_b200 import Engine, synthetic

    kind, N, S, B = "emission_absorption_physical", 8, 4096, 2048
    eng = Engine(kind, N, lorentzian=True, prior_fwhm_L=1.0, prior_amplitude=1e21)
    v = np.linspace(-100, 100, S)
    eng.set_spectra(emission=dict(spectral=v, brightness=np.zeros(S), noise=0.1, basis=np.full((1, S), 0.1)),
                    absorption=dict(spectral=v, brightness=np.zeros(S), noise=0.01, basis=np.full((1, S), 0.01)))
    th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal"]
    rng = np.random.default_rng(1234)
    theta = torch.from_numpy(eng.pack(synthetic.draw_free(kind, N, B, rng, lorentzian=True, thermal_fn=th))).cuda()
    ce = torch.from_numpy(0.3 * rng.standard_normal((B, 1))).cuda()
    ca = torch.from_numpy(0.3 * rng.standard_normal((B, 1))).cuda()
    e64 = torch.empty((B, S), dtype=torch.float64, device="cuda"); a64 = torch.empty_like(e64)
    e32 = torch.empty((B, S), dtype=torch.float32, device="cuda"); a32 = torch.empty_like(e32)
    eng.spectra_dev(B, theta, e64, a64, ce, ca)
    eng.spectra_f32_dev(B, theta, e32, a32, ce, ca)
    eng.synchronize()
    assert_close(e32.cpu().numpy(), e64.cpu().numpy(), rtol=RTOL32, what="config 5 emission f32 vs f64")
    assert_close(a32.cpu().numpy(), a64.cpu().numpy(), rtol=RTOL32, what="config 5 absorption f32 vs f64")
    assert int(torch.isfinite(e64).all(dim=1).sum()) > B // 4  # 8-cloud prior draws: about half are NaN in the reference too
    eng.close()
