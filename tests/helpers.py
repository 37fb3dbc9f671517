"""Shared test helpers: seeded cases, oracle evaluation (torch-CPU fp64 autograd) and
the comparison rule of SURVEY.md section 8c.  Only tests import the oracle."""

import numpy as np
import torch

from oracle import models_ref as mr

RTOL64 = 1e-10  # BASELINE.json north_star: 1e-10 relative in the fp64 path


def assert_close(x, ref, rtol=RTOL64, what="", axis=-1):
    """abs(x-ref) <= rtol*abs(ref) + atol with atol = rtol*max|ref| over `axis`
    (the output vector); NaN/inf patterns must match exactly."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert x.shape == ref.shape, f"{what}: shape {x.shape} vs {ref.shape}"
    bad_ref = ~np.isfinite(ref)
    bad_x = ~np.isfinite(x)
    if not np.array_equal(bad_ref, bad_x):
        i = tuple(np.argwhere(bad_ref != bad_x)[0])
        raise AssertionError(f"{what}: non-finite pattern differs at {i}: x={x[i]!r}, ref={ref[i]!r}")
    if ref.ndim == 0 or ref.size == 0:
        scale = np.abs(ref)
    else:
        scale = np.nanmax(np.where(bad_ref, 0.0, np.abs(ref)), axis=axis, keepdims=True)
    err = np.where(bad_ref, 0.0, np.abs(x - ref))
    tol = rtol * np.where(bad_ref, 0.0, np.abs(ref)) + rtol * scale
    worst = np.max(err - tol) if err.size else 0.0
    if worst > 0:
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(
            f"{what}: |x-ref|={err[i]:.3e} > tol={tol[i]:.3e} at {i} (x={x[i]!r}, ref={ref[i]!r}, "
            f"rel={err[i] / max(abs(ref[i]), 1e-300):.3e})"
        )


def draw_free_oracle(spec, B, rng):
    """Prior draws of the free RVs using the oracle for the derived quantity the
    weight prior needs (CPU-only twin of caribou_hi_b200.synthetic.draw_free)."""
    from caribou_hi_b200 import synthetic

    key = "fwhm2_thermal" if mr.is_physical(spec) else "tkin"

    def thermal(trial):
        return mr.cloud_inputs(spec, trial)[key]

    return synthetic.draw_free(
        spec["kind"], spec["n_clouds"], B, rng,
        lorentzian=spec["prior_fwhm_L"] is not None,
        free_weight=spec["prior_sigma_log10_NHI"] is not None,
        sigma_log10_NHI=spec["prior_sigma_log10_NHI"] or 0.5,
        prior_nth_fwhm_1pc=tuple(spec["prior_nth_fwhm_1pc"]),
        thermal_fn=thermal,
    )


def make_case(kind, n_clouds, S_e, S_a, B, seed, lorentzian=False, baseline_degree=0, weight=True,
              half_width=60.0, shared_axis=False, **spec_kwargs):
    """A seeded test case: oracle spec, oracle data (truth at one prior draw + noise),
    B prior draws of the free RVs and baseline coefficients."""
    rng = np.random.default_rng(seed)
    kw = dict(spec_kwargs)
    if lorentzian:
        kw.setdefault("prior_fwhm_L", 1.0)
    if kind == "emission_absorption":
        kw.setdefault("prior_sigma_log10_NHI", 0.5 if weight else None)
    spec = mr.make_spec(kind, n_clouds, **kw)
    free = draw_free_oracle(spec, B, rng)
    baseline = {}
    axes = {}
    if mr.has_emission(spec):
        axes["emission"] = (np.linspace(-half_width, half_width, S_e), 0.1)
    if mr.has_absorption(spec):
        if shared_axis:  # both spectra on one velocity grid: exercises the fused kernel
            axes["absorption"] = (np.linspace(-half_width, half_width, S_e), 0.01)
        else:
            axes["absorption"] = (np.linspace(-half_width / 2, half_width / 2, S_a), 0.01)
    for key in axes:
        baseline[f"baseline_{key}_norm"] = 0.3 * rng.standard_normal(size=(B, baseline_degree + 1))
    # truth: forward model of a moderate, well-behaved draw + noise
    truth_free = {k: np.array(v[:1]) for k, v in draw_free_oracle(spec, 1, np.random.default_rng(seed + 1)).items()}
    dummy = {k: mr.OracleData(ax, np.zeros_like(ax), s) for k, (ax, s) in axes.items()}
    truth = mr.predict(spec, dummy, truth_free)
    data = {}
    for key, (ax, s) in axes.items():
        y = np.nan_to_num(truth[key][0]) + s * rng.standard_normal(size=ax.shape)
        data[key] = mr.OracleData(ax, y, s)
    return spec, data, free, baseline


def oracle_logp_grad(spec, data, free, baseline):
    """logp [B] and gradients (dict keyed by RV name) by torch autograd."""
    ft = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in free.items()}
    bt = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in baseline.items()}
    lp = mr.logp(spec, data, ft, bt)
    # each draw only depends on its own row, so one backward of the sum gives per-draw grads
    torch.nan_to_num(lp, nan=0.0, posinf=0.0, neginf=0.0).sum().backward()
    grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros_like(free[k])) for k, v in ft.items()}
    grads.update({k: (v.grad.numpy() if v.grad is not None else np.zeros_like(baseline[k])) for k, v in bt.items()})
    return lp.detach().numpy(), grads


def engine_for(spec, data, baseline_degree=0, device=0):
    """A caribou_hi_b200.Engine configured exactly like the oracle spec."""
    from caribou_hi_b200 import Engine

    amp = {
        "absorption": "prior_tau_total",
        "emission": "prior_TB_fwhm",
        "emission_absorption": "prior_TB_fwhm",
        "absorption_physical": "prior_NHI_fwhm2_thermal",
        "emission_physical": "prior_ff_NHI",
        "emission_absorption_physical": "prior_ff_NHI",
    }[spec["kind"]]
    eng = Engine(
        spec["kind"], spec["n_clouds"], device=device,
        lorentzian=spec["prior_fwhm_L"] is not None,
        free_weight=spec["prior_sigma_log10_NHI"] is not None,
        bg_temp=spec["bg_temp"], depth_nth_fwhm_power=spec["depth_nth_fwhm_power"],
        prior_fwhm2=spec["prior_fwhm2"], prior_velocity=spec["prior_velocity"],
        prior_log10_nHI=spec["prior_log10_nHI"], prior_log10_n_alpha=spec["prior_log10_n_alpha"],
        prior_fwhm_L=spec["prior_fwhm_L"] or 0.0, prior_amplitude=spec[amp],
    )
    sets = {}
    for key, d in data.items():
        n_terms = baseline_degree + 1
        basis = np.stack([d.spectral_norm**i / (i + 1.0) ** i * d._brightness_scale for i in range(n_terms)])
        sets[key] = dict(spectral=d.spectral, brightness=d.brightness, noise=d.noise, basis=basis,
                         offset=d._brightness_offset)
    eng.set_spectra(emission=sets.get("emission"), absorption=sets.get("absorption"))
    return eng
