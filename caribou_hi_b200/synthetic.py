"""Seeded synthetic inputs of the shapes BASELINE.json names: free-RV draws from the
reference's default priors and spectral axes/noise like the reference's tutorial
(docs/source/notebooks/optimization.ipynb cell 3).  Pure NumPy host code; no model
arithmetic happens here (derived quantities come from the CUDA engine).
"""

import numpy as np

from . import layout


def draw_free(kind, n_clouds, B, rng, *, lorentzian=False, free_weight=True, sigma_log10_NHI=0.5,
              prior_nth_fwhm_1pc=(1.75, 0.25), thermal_fn=None):
    """Draw the model's free RVs from the reference's default prior distributions:

    * ``fwhm2_norm ~ ChiSquared(1)`` (hi_model.py:95), ``velocity_norm, log10_nHI_norm,
      log10_n_alpha_norm ~ Normal(0,1)`` (hi_model.py:99-129)
    * ``nth_fwhm_1pc ~ TruncatedNormal(mu, sigma, lower=0)`` (hi_physical_model.py:138)
    * ``fwhm_L_norm, *_norm amplitudes ~ HalfNormal(1)``; thermal fractions ``~ Beta(2,2)``;
      ``filling_factor_norm ~ Uniform(0,1)``
    * ``log10_wt_* ~ Normal(-ln10 sigma^2/2 - log10(T), sigma)`` with ``T`` = tkin or
      fwhm2_thermal from ``thermal_fn(free) -> [B, N]`` (emission_absorption_model.py:50-55,
      emission_absorption_physical_model.py:43-52)

    Returns a dict of ``[B, N]`` arrays (``[B, 1]`` for the physical models' shared
    ``fwhm_L_norm``).
    """
    N = n_clouds
    names = layout.slot_names(kind, lorentzian, free_weight)
    free = {}
    for name in names:
        if name == "fwhm2_norm":
            free[name] = rng.chisquare(1.0, size=(B, N))
        elif name in ("velocity_norm", "log10_nHI_norm", "log10_n_alpha_norm"):
            free[name] = rng.standard_normal(size=(B, N))
        elif name == "nth_fwhm_1pc":
            mu, sig = prior_nth_fwhm_1pc
            x = mu + sig * rng.standard_normal(size=(B, N))
            bad = x <= 0.0
            while bad.any():  # rejection for the lower=0 truncation (7 sigma away by default)
                x[bad] = mu + sig * rng.standard_normal(size=int(bad.sum()))
                bad = x <= 0.0
            free[name] = x
        elif name == "fwhm_L_norm":
            shape = (B, 1) if layout.is_physical(kind) else (B, N)
            free[name] = np.abs(rng.standard_normal(size=shape))
        elif name in ("tau_total_norm", "TB_fwhm_norm", "NHI_fwhm2_thermal_norm", "ff_NHI_norm"):
            free[name] = np.abs(rng.standard_normal(size=(B, N)))
        elif name in ("tkin_factor", "tkin_factor_norm", "fwhm2_thermal_fraction", "fwhm2_thermal_fraction_norm"):
            free[name] = rng.beta(2.0, 2.0, size=(B, N))
        elif name == "filling_factor_norm":
            free[name] = rng.uniform(0.0, 1.0, size=(B, N))
        elif name in ("log10_wt_ff_tkin", "log10_wt_ff_fwhm2_thermal"):
            pass  # needs a derived quantity; drawn below
        else:  # pragma: no cover
            raise KeyError(name)
    for name in ("log10_wt_ff_tkin", "log10_wt_ff_fwhm2_thermal"):
        if name in names:
            z = rng.standard_normal(size=(B, N))
            if thermal_fn is None:
                raise ValueError(f"{name} needs thermal_fn to evaluate its prior mean")
            trial = dict(free)
            trial[name] = np.zeros((B, N))
            T = np.asarray(thermal_fn(trial), dtype=np.float64)
            mu = -np.log(10.0) * sigma_log10_NHI**2.0 / 2.0 - np.log10(T)
            free[name] = mu + sigma_log10_NHI * z
    return free


def spectral_axes(S_e, S_a, half_width=100.0):
    """Uniform velocity axes (km/s), like the tutorial's ``np.linspace``."""
    return np.linspace(-half_width, half_width, S_e) if S_e else None, (
        np.linspace(-half_width, half_width, S_a) if S_a else None
    )


NOISE_EMISSION = 0.1  # K      (optimization.ipynb cell 3)
NOISE_ABSORPTION = 0.01  # 1-exp(-tau)
