"""Arbitration of the 1e-10 tolerance with 60-digit arithmetic (mpmath): single gradient
components are sums of large cancelling channel terms, so fp64 autograd itself sits ~1e-10
(relative to the component) away from the exact value, while it is exact to ~1e-13 relative
to the draw's gradient infinity norm -- which is why the parity tests scale atol by that norm."""

import mpmath as mp
import numpy as np

from helpers import make_case, oracle_logp_grad
from oracle import models_ref as mr
from oracle.backend import MpmathBackend as MB


def test_autograd_vs_high_precision():
    mp.mp.dps = 60
    spec, data, free, baseline = make_case("emission_absorption_physical", 1, 257, 131, 16, seed=101,
                                           lorentzian=False, baseline_degree=1)
    lp, g = oracle_logp_grad(spec, data, free, baseline)
    b, name, n = 3, "filling_factor_norm", 0

    def f(delta):
        fr = {k: MB.asarray(v[b : b + 1]) for k, v in free.items()}
        bl = {k: MB.asarray(v[b : b + 1]) for k, v in baseline.items()}
        fr[name] = fr[name].copy()
        fr[name][0, n] = fr[name][0, n] + delta
        return mr.logp(spec, data, fr, bl)[0]

    h = mp.mpf(10) ** (-25)
    exact = (f(h) - f(-h)) / (2 * h)
    auto = g[name][b, n]
    rel_component = abs(float((mp.mpf(auto) - exact) / exact))
    norm = max(np.abs(v[b]).max() for v in g.values())
    rel_norm = abs(float(mp.mpf(auto) - exact)) / norm
    # the high-precision logp agrees with fp64 to rounding
    assert abs(float(f(0) - mp.mpf(lp[b])) / lp[b]) < 1e-14
    assert rel_component < 1e-8  # poorly conditioned component: ~1.4e-10 in practice
    assert rel_norm < 1e-13  # well inside the 1e-10 budget once scaled by the gradient norm


def test_extreme_draw():
    """Draw 50 (finite ones) of tests/test_gpu_parity.py::test_branch_edges_of_the_parameter_maps
    [absorption_physical], amplitude swept over six decades: the oracle's fp64 autograd is 6e-10
    away from the 60-digit derivative; the value the CUDA path returned (recorded below) is 1.3e-10
    away.  The disagreement the GPU test tolerates at 2e-9 is the checker's rounding."""
    mp.mp.dps = 60
    kind, B, N = "absorption_physical", 768, 3
    spec, data, free, baseline = make_case(kind, N, 96, 64, B, seed=909, lorentzian=True)
    rng = np.random.default_rng(910)
    free["NHI_fwhm2_thermal_norm"] = free["NHI_fwhm2_thermal_norm"] * 10 ** rng.uniform(-3, 3, size=(B, 1))
    free["fwhm2_norm"] = free["fwhm2_norm"] * 10 ** rng.uniform(-2, 1, size=(B, 1))
    free["fwhm_L_norm"][B // 2 :: 7] = 0.0
    lp, g = oracle_logp_grad(spec, data, free, baseline)
    b = np.where(np.isfinite(lp))[0][50]
    name, n = "fwhm2_norm", 0

    def f(delta):
        fr = {k: MB.asarray(v[b : b + 1]) for k, v in free.items()}
        bl = {k: MB.asarray(v[b : b + 1]) for k, v in baseline.items()}
        fr[name] = fr[name].copy()
        fr[name][0, n] = fr[name][0, n] + delta
        return mr.logp(spec, data, fr, bl)[0]

    h = mp.mpf(10) ** (-25)
    exact = (f(h) - f(-h)) / (2 * h)
    cuda_value = -44254.205033823055  # from the B200 run of the GPU test
    err_auto = abs(float((mp.mpf(float(g[name][b, n])) - exact) / exact))
    err_cuda = abs(float((mp.mpf(cuda_value) - exact) / exact))
    assert 1e-10 < err_auto < 2e-9
    assert err_cuda < 2e-10 and err_cuda < err_auto
