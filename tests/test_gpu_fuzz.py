"""Seeded random configurations through every kernel family -- model kind, 1..16 clouds (padded
instantiations above 8), ragged channel counts, shared / different axes (fused / single-dataset
kernels), Gaussian / pseudo-Voigt, baseline degree 0..3 (simple-baseline and general-baseline
instantiations), one or several sightlines (L1-cached / bulk-async staged kernels), batch sizes on
both sides of the chunking and split-epilogue thresholds, slot and packed-row entry points --
against the oracle at the 1e-10 bar."""

import numpy as np
import pytest

from helpers import assert_close, assert_grad_close, make_case, oracle_logp_grad
from oracle import models_ref as mr

pytestmark = pytest.mark.gpu

KINDS = list(mr.KINDS)


def _config(i):
    rng = np.random.default_rng(7000 + i)
    kind = KINDS[i % len(KINDS)]
    return dict(
        kind=kind,
        N=int(rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 12, 14, 16])),
        S_e=int(rng.choice([2, 31, 32, 33, 64, 97, 150, 257])),
        S_a=int(rng.choice([2, 17, 32, 65, 128, 190])),
        shared=bool(rng.integers(2)),
        lor=bool(rng.integers(2)),
        deg=int(rng.choice([0, 0, 1, 2, 3])),
        B=int(rng.choice([1, 7, 64, 300])),
        n_sl=int(rng.choice([1, 1, 3])),
        rows=bool(rng.integers(2)),
        seed=int(rng.integers(1 << 30)),
    )


@pytest.mark.parametrize("i", range(72))
def test_random_configuration(cuda_device, i):
    from caribou_hi_b200 import Engine

    c = _config(i)
    spec, data, free, baseline = make_case(c["kind"], c["N"], c["S_e"], c["S_a"], c["B"], c["seed"], lorentzian=c["lor"],
                                           baseline_degree=c["deg"], shared_axis=c["shared"])
    rng = np.random.default_rng(c["seed"] + 1)
    n_sl, B = c["n_sl"], c["B"]
    # per-sightline spectra: the case's data plus independent perturbations
    sets, per_sl = {}, []
    for key, d in data.items():
        y = d.brightness[None, :] + (0.0 if n_sl == 1 else 3.0 * d.noise[None, :] * rng.standard_normal((n_sl, d.spectral.size)))
        n_terms = c["deg"] + 1
        basis = np.stack([d.spectral_norm**k / (k + 1.0) ** k * d._brightness_scale for k in range(n_terms)])
        sets[key] = dict(spectral=d.spectral, brightness=y if n_sl > 1 else y[0], noise=d.noise, basis=basis,
                         offset=d._brightness_offset)
    for j in range(n_sl):
        dj = {}
        for key, d in data.items():
            o = mr.OracleData(d.spectral, np.atleast_2d(sets[key]["brightness"])[j], d.noise)
            # keep the case's normalisation (the engine was given ONE basis / offset per dataset)
            o._brightness_offset, o._brightness_scale = d._brightness_offset, d._brightness_scale
            dj[key] = o
        per_sl.append(dj)
    amp = {"absorption": "prior_tau_total", "emission": "prior_TB_fwhm", "emission_absorption": "prior_TB_fwhm",
           "absorption_physical": "prior_NHI_fwhm2_thermal", "emission_physical": "prior_ff_NHI",
           "emission_absorption_physical": "prior_ff_NHI"}[c["kind"]]
    eng = Engine(c["kind"], c["N"], lorentzian=c["lor"], free_weight=spec["prior_sigma_log10_NHI"] is not None,
                 bg_temp=spec["bg_temp"], depth_nth_fwhm_power=spec["depth_nth_fwhm_power"],
                 prior_fwhm2=spec["prior_fwhm2"], prior_velocity=spec["prior_velocity"],
                 prior_log10_nHI=spec["prior_log10_nHI"], prior_log10_n_alpha=spec["prior_log10_n_alpha"],
                 prior_fwhm_L=spec["prior_fwhm_L"] or 0.0, prior_amplitude=spec[amp])
    eng.set_spectra(emission=sets.get("emission"), absorption=sets.get("absorption"), n_sightlines=n_sl)
    sl = rng.integers(0, n_sl, size=B).astype(np.int32) if n_sl > 1 else None
    ce, ca = baseline.get("baseline_emission_norm"), baseline.get("baseline_absorption_norm")
    if c["rows"]:
        lp, grad_r, gce, gca = eng.logp_grad_rows(eng.pack_rows(free), ce, ca, sightline=sl)
        got = eng.unpack_rows(grad_r)
    else:
        lp, grad, gce, gca = eng.logp_grad(eng.pack(free), ce, ca, sightline=sl)
        got = eng.unpack_grad(grad)
    names = sorted(got)
    for j in range(n_sl):
        sel = np.arange(B) if sl is None else np.where(sl == j)[0]
        if sel.size == 0:
            continue
        ref_lp, ref_g = oracle_logp_grad(spec, per_sl[j], {k: v[sel] for k, v in free.items()},
                                         {k: v[sel] for k, v in baseline.items()})
        assert_close(lp[sel], ref_lp, what=f"{c} logp", axis=None)
        fin = np.isfinite(ref_lp)
        g_sel = {n: got[n][sel] for n in names}
        if gce is not None:
            g_sel["baseline_emission_norm"] = gce[sel]
        if gca is not None:
            g_sel["baseline_absorption_norm"] = gca[sel]
        # per RV vector; components that miss 1e-10 against autograd (tiny chi-square draws give kinetic
        # temperatures of 1e-5 K where fwhm2 - fwhm2_thermal cancels) go to the 60-digit arbiter
        case = (spec, per_sl[j], {k: v[sel] for k, v in free.items()}, {k: v[sel] for k, v in baseline.items()})
        assert_grad_close(g_sel, ref_g, fin, case=case, what=f"{c} gradient", max_arbitrations=60)
    if n_sl == 1:  # predicted spectra (fused forward kernel on a shared axis, one kernel per dataset otherwise)
        ref_mu = mr.predict(spec, per_sl[0], free, baseline)
        mu_e, mu_a = eng.spectra(eng.pack(free), ce, ca)
        if mu_e is not None:
            assert_close(mu_e, ref_mu["emission"], what=f"{c} mu_e")
        if mu_a is not None:
            assert_close(mu_a, ref_mu["absorption"], what=f"{c} mu_a")
    eng.close()
