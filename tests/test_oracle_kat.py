"""Pins the CPU oracle against every known-answer value the reference tree holds for
this path: tests/test_physics.py and the printed outputs of
docs/source/notebooks/optimization.ipynb cell 5."""

import numpy as np
import pytest

from oracle import physics_ref as ph


def test_gaussian_kat():  # reference tests/test_physics.py:15-21
    val = np.exp(-4.0 * np.log(2.0)) * np.sqrt(4.0 * np.log(2.0) / np.pi)
    assert ph.gaussian(1.0, 0.0, 1.0) == pytest.approx(val)


def test_lorentzian_kat():  # tests/test_physics.py:24-30
    assert ph.lorentzian(1.0, 0.0, 1.0) == pytest.approx(1.0 / (2.0 * np.pi) / 1.25)


def test_spin_temp_kat():  # tests/test_physics.py:33-43
    assert ph.calc_spin_temp(1000.0, 1000.0, 1.0e-6) == pytest.approx(1000.0, abs=1.0)
    assert ph.calc_spin_temp(8000.0, 1.0, 1.0e-6) < 8000.0


def test_scalar_helpers_kat():  # tests/test_physics.py:46-99
    assert ph.calc_thermal_fwhm2(1000.0) == pytest.approx(0.2139**2.0 * 1000.0)
    assert ph.calc_nonthermal_fwhm(2.0, 2.0, 2.0) == pytest.approx(8.0)
    assert ph.calc_depth_nonthermal(8.0, 2.0, 2.0) == pytest.approx(2.0)
    assert ph.calc_log10_column_density(1.0, 1.0) == pytest.approx(np.log10(1.82243e18))
    assert ph.calc_log10_density(18.489351, 0.0) == pytest.approx(0.0)
    assert ph.calc_tau_total(1.82243e18, 1.0) == pytest.approx(1.0)
    assert ph.calc_log10_nonthermal_pressure(0.0, 1.0) == pytest.approx(np.log10(671.8))


def test_profile_and_transfer_shapes():  # tests/test_physics.py:102-129
    v = np.linspace(-100.0, 100.0, 101)
    prof = ph.calc_pseudo_voigt(v, np.array([-5.0, 5.0]), np.array([100.0, 200.0]), 0.0)
    assert prof.shape == (101, 2) and not np.any(np.isnan(prof))
    tau = np.array([5.0, 3.0]) * prof
    tb = ph.radiative_transfer(tau, np.array([1000.0, 5000.0]), 1.0, 2.7)
    assert tb.shape == (101,) and not np.any(np.isnan(tb))
    # profile integrates to ~1 per cloud (normalised line profile)
    assert np.allclose(prof.sum(axis=0) * 2.0, 1.0, rtol=1e-3)


def test_notebook_cell5_golden():
    """optimization.ipynb cell 5 printed outputs, all printed digits."""
    tkin = np.array([50.0, 300.0, 5000.0])
    log10_nHI = np.array([1.25, 0.0, -0.5])
    log10_n_alpha = np.array([-5.0, -6.0, -7.0])
    depth = np.array([5.0, 25.0, 300.0])
    nth = np.array([2.0, 1.75, 1.5])
    ff = np.array([0.2, 0.8, 1.0])
    wt = np.array([0.9, 0.8, 1.0])
    tspin = ph.calc_spin_temp(tkin, 10.0**log10_nHI, 10.0**log10_n_alpha)
    np.testing.assert_allclose(tspin, [49.99731723, 297.73994712, 2378.08874354], rtol=0, atol=5e-9)
    f2n = ph.calc_nonthermal_fwhm(depth, nth, 1 / 3) ** 2.0
    np.testing.assert_allclose(f2n, [11.69607095, 26.18400668, 100.8316068], rtol=0, atol=5e-9)
    f2t = ph.calc_thermal_fwhm2(tkin)
    np.testing.assert_allclose(f2t, [2.2876605, 13.725963, 228.76605], rtol=0, atol=5e-9)
    fwhm2 = f2n + f2t
    np.testing.assert_allclose(fwhm2, [13.98373145, 39.90996968, 329.5976568], rtol=0, atol=5e-9)
    np.testing.assert_allclose(f2t / fwhm2, [0.16359442, 0.34392316, 0.69407669], rtol=0, atol=5e-9)
    np.testing.assert_allclose(np.log10(tkin) + log10_nHI, [2.94897, 2.47712125, 3.19897], rtol=0, atol=5e-9)
    log10_NHI = log10_nHI + np.log10(depth) + 18.489351
    np.testing.assert_allclose(log10_NHI, [20.438321, 19.88729101, 20.46647225], rtol=0, atol=5e-9)
    tau_total = ph.calc_tau_total(10.0**log10_NHI, tspin)
    np.testing.assert_allclose(tau_total, [3.01108799, 0.14216839, 0.06754502], rtol=0, atol=5e-9)
    np.testing.assert_allclose(np.log10(wt / ff / f2t), [0.29382094, -1.13754282, -2.35939157], rtol=0, atol=5e-9)


def test_notebook_cell5_through_physical_model_map():
    """The same golden numbers recovered through the restated EmissionAbsorptionPhysical
    parameter map (emission_physical_model.py:59-129 + hi_physical_model.py:153-197),
    i.e. the map inverts the notebook's simulation inputs."""
    from oracle import models_ref as mr

    spec = mr.make_spec("emission_absorption_physical", 3)
    fwhm2 = np.array([13.98373145, 39.90996968, 329.5976568])
    tkin = np.array([50.0, 300.0, 5000.0])
    f2t = 0.2139**2 * tkin
    fwhm2 = f2t + ph.calc_nonthermal_fwhm(np.array([5.0, 25.0, 300.0]), np.array([2.0, 1.75, 1.5]), 1 / 3) ** 2
    ff = np.array([0.2, 0.8, 1.0])
    log10_NHI = np.array([1.25, 0.0, -0.5]) + np.log10([5.0, 25.0, 300.0]) + 18.489351
    ff_NHI = ff * 10.0**log10_NHI
    const_G = np.sqrt(2.0 * np.pi) / (2.0 * np.sqrt(2.0 * np.log(2.0)))
    tkin_min = (ff_NHI / 1.82243e18) / (const_G * np.sqrt(fwhm2))
    frac_min = 0.2139**2 * tkin_min / fwhm2
    frac = f2t / fwhm2
    ff_min = tkin_min / tkin
    free = {
        "fwhm2_norm": fwhm2 / 500.0,
        "velocity_norm": np.array([-0.5, 0.5, 1.0]),
        "log10_n_alpha_norm": (np.array([-5.0, -6.0, -7.0]) + 6.0) / 2.0,
        "nth_fwhm_1pc": np.array([2.0, 1.75, 1.5]),
        "ff_NHI_norm": ff_NHI / 1e21,
        "fwhm2_thermal_fraction_norm": (frac - frac_min) / (1 - frac_min),
        "filling_factor_norm": np.where(ff_min > 1, 0.5, (ff - ff_min) / (1 - ff_min)),
        "log10_wt_ff_fwhm2_thermal": np.array([0.29382094, -1.13754282, -2.35939157]),
    }
    det = mr.cloud_inputs(spec, free)
    np.testing.assert_allclose(det["tkin"], tkin, rtol=1e-12)
    np.testing.assert_allclose(det["tspin"][:2], [49.99731723, 297.73994712], rtol=0, atol=5e-8)
    np.testing.assert_allclose(det["tau_total"][:2], [3.01108799, 0.14216839], rtol=0, atol=5e-8)
    np.testing.assert_allclose(det["depth"][:2], [5.0, 25.0], rtol=1e-9)
    np.testing.assert_allclose(det["absorption_weight"][:2], [0.9, 0.8], rtol=1e-7)
