"""Predictive draws of the observed nodes (chi_predictive): y = mu + sigma z, z from the counter-based
generator documented in include/chi.h.  The generator is restated here in NumPy (splitmix64 finaliser
on (seed, draw, dataset, channel pair), Box-Muller), so the check is exact up to libm rounding, and
the statistics of z are checked on top."""

import numpy as np
import pytest

from helpers import assert_close, engine_for, make_case
from oracle import models_ref as mr

pytestmark = pytest.mark.gpu

M64 = (1 << 64) - 1
GOLD = 0x9E3779B97F4A7C15


def mix64(z):
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def ref_normals(seed, draw, dataset, S, f32=False):
    """z[S] of one (draw, dataset) by the documented recipe (Python integers: exact)."""
    base = mix64((seed + GOLD * (2 * draw + dataset + 1)) & M64)
    z = np.empty(2 * ((S + 1) // 2))
    for p in range((S + 1) // 2):
        h1 = mix64((base + GOLD * (2 * p + 1)) & M64)
        if f32:
            u1 = ((h1 >> 40) + 0.5) / 16777216.0
            u2 = (((h1 >> 16) & 0xFFFFFF) + 0.5) / 16777216.0
        else:
            h2 = mix64((base + GOLD * (2 * p + 2)) & M64)
            u1 = ((h1 >> 11) + 0.5) / 9007199254740992.0
            u2 = ((h2 >> 11) + 0.5) / 9007199254740992.0
        r = np.sqrt(-2.0 * np.log(u1))
        z[2 * p] = r * np.cos(2 * np.pi * u2)
        z[2 * p + 1] = r * np.sin(2 * np.pi * u2)
    return z[:S]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shared", [False, True])
def test_predictive_draws_match_the_restated_generator(cuda_device, dtype, shared):
    B, S_e, S_a, seed, off = 6, 101, 64, 20240607, 1000
    spec, data, free, baseline = make_case("emission_absorption", 2, S_e, S_a, B, seed=5, lorentzian=True,
                                           baseline_degree=1, shared_axis=shared)
    if shared:
        S_a = S_e
    # per-channel noise, so that a wrong sigma index shows
    rng = np.random.default_rng(3)
    for key, S in (("emission", S_e), ("absorption", S_a)):
        d = data[key]
        data[key] = mr.OracleData(d.spectral, d.brightness, d.noise * rng.uniform(0.5, 2.0, S))
    eng = engine_for(spec, data, 1)
    theta = eng.pack(free)
    ce, ca = baseline["baseline_emission_norm"], baseline["baseline_absorption_norm"]
    mu_e, mu_a = eng.spectra(theta, ce, ca)
    y_e, y_a = eng.predictive(theta, ce, ca, seed=seed, draw_offset=off, dtype=dtype)
    assert y_e.dtype == dtype and y_a.dtype == dtype
    f32 = dtype == np.float32
    rtol = 1e-5 if f32 else 1e-10
    for y, mu, key, ds, S in ((y_e, mu_e, "emission", 0, S_e), (y_a, mu_a, "absorption", 1, S_a)):
        ref = np.stack([mu[b] + data[key].noise * ref_normals(seed, off + b, ds, S, f32) for b in range(B)])
        assert_close(y, ref, rtol=rtol, what=f"predictive {key}")
    # a shard that passes its offset reproduces its rows of the big call bit for bit; another seed does not
    y2_e, y2_a = eng.predictive(theta[2:5], ce[2:5], ca[2:5], seed=seed, draw_offset=off + 2, dtype=dtype)
    assert np.array_equal(y2_e, y_e[2:5], equal_nan=True) and np.array_equal(y2_a, y_a[2:5], equal_nan=True)
    y3_e, _ = eng.predictive(theta, ce, ca, seed=seed + 1, draw_offset=off, dtype=dtype)
    fin = np.isfinite(y_e).all(axis=1)
    assert fin.any() and not np.any(y3_e[fin] == y_e[fin])
    eng.close()


def test_predictive_noise_statistics_sightlines_and_device_form(cuda_device):
    """z = (y - mu) / sigma over 2e6 samples: moments of N(0, 1), no correlation between neighbouring
    channels, draws or datasets; per-draw sightline noise; device-pointer form == host form."""
    import torch
    from caribou_hi_b200 import Engine

    B, N, S, n_sl = 1000, 2, 1000, 3
    rng = np.random.default_rng(11)
    v = np.linspace(-50, 50, S)
    sig_e = rng.uniform(0.05, 0.3, (n_sl, S)); sig_a = rng.uniform(0.005, 0.03, (n_sl, S))
    eng = Engine("physics", N, lorentzian=False)
    eng.set_spectra(emission=dict(spectral=v, brightness=np.zeros((n_sl, S)), noise=sig_e),
                    absorption=dict(spectral=v, brightness=np.zeros((n_sl, S)), noise=sig_a), n_sightlines=n_sl)
    inp = {"velocity": rng.normal(0, 10, (B, N)), "fwhm2": rng.uniform(20, 200, (B, N)), "fwhm_L": np.zeros((B, N)),
           "tau_total": rng.uniform(0.1, 3, (B, N)), "tspin": rng.uniform(30, 300, (B, N)),
           "filling_factor": rng.uniform(0.3, 1, (B, N)), "absorption_weight": np.ones((B, N))}
    theta = eng.pack(inp)
    sl = rng.integers(0, n_sl, B).astype(np.int32)
    mu_e, mu_a = eng.spectra(theta)
    y_e, y_a = eng.predictive(theta, seed=7, sightline=sl)
    z_e = (y_e - mu_e) / sig_e[sl]
    z_a = (y_a - mu_a) / sig_a[sl]
    for z in (z_e, z_a):
        n = z.size
        assert abs(z.mean()) < 5 / np.sqrt(n)
        assert abs(z.var() - 1) < 5 * np.sqrt(2 / n)
        assert abs((z**3).mean()) < 5 * np.sqrt(15 / n)
        assert abs((z**4).mean() - 3) < 5 * np.sqrt(96 / n)
        assert abs((z[:, 1:] * z[:, :-1]).mean()) < 5 / np.sqrt(n)   # neighbouring channels (Box-Muller pair and next pair)
        assert abs((z[1:] * z[:-1]).mean()) < 5 / np.sqrt(n)         # neighbouring draws
        assert np.abs(z).max() < 7
    assert abs((z_e * z_a).mean()) < 5 / np.sqrt(z_e.size)           # the two datasets
    # device-pointer form, fp64 and fp32
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    d_e = torch.empty((B, S), dtype=torch.float64, device="cuda"); d_a = torch.empty_like(d_e)
    eng.predictive_dev(B, t(theta), d_e, d_a, seed=7, sightline=t(sl))
    eng.synchronize()
    assert np.array_equal(d_e.cpu().numpy(), y_e) and np.array_equal(d_a.cpu().numpy(), y_a)
    f_e = torch.empty((B, S), dtype=torch.float32, device="cuda"); f_a = torch.empty_like(f_e)
    eng.predictive_dev(B, t(theta), f_e, f_a, seed=7, sightline=t(sl), f32=True)
    eng.synchronize()
    h_e, h_a = eng.predictive(theta, seed=7, sightline=sl, dtype=np.float32)
    assert np.array_equal(f_e.cpu().numpy(), h_e) and np.array_equal(f_a.cpu().numpy(), h_a)
    z32 = (h_e.astype(np.float64) - mu_e) / sig_e[sl]
    assert abs(z32.mean()) < 5 / np.sqrt(z32.size) and abs(z32.var() - 1) < 5 * np.sqrt(2 / z32.size)
    with pytest.raises(Exception):
        eng.predictive(theta, seed=7, sightline=np.full(B, n_sl, dtype=np.int32))
    eng.close()


def test_model_level_sample_predictive(cuda_device):
    from caribou_hi_b200 import EmissionAbsorptionModel, SpecData

    rng = np.random.default_rng(2)
    v = np.linspace(-40, 40, 200)
    data = {"emission": SpecData(v, rng.normal(5, 1, 200), 0.2), "absorption": SpecData(v, rng.normal(0.2, 0.05, 200), 0.02)}
    model = EmissionAbsorptionModel(data, n_clouds=2, baseline_degree=0)
    model.add_priors()
    model.add_likelihood()
    draws = model.sample_prior(500, np.random.default_rng(4))
    mu = model.predict(draws)
    y = model.sample_predictive(draws, seed=99)
    for key, s in (("emission", 0.2), ("absorption", 0.02)):
        fin = np.isfinite(mu[key]).all(axis=1)
        z = (y[key][fin] - mu[key][fin]) / s
        assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.var() - 1) < 5 * np.sqrt(2 / z.size)
    y2 = model.sample_predictive(draws, seed=99)
    assert all(np.array_equal(y[k], y2[k], equal_nan=True) for k in y)
