"""Oracle restatement of the prior terms and default transforms that PyMC adds to the
likelihood when it builds ``model.logp`` in the unconstrained space NUTS works in
(test infrastructure).

PyMC is third-party (pymc>=5; the notebook ran 5.22.0) and NOT installable here, so the
densities, transforms and Jacobians below are restated from its published definitions and are
UNVERIFIED against a running PyMC ("parity unpinned"):

  distribution (reference file:line)                      value variable        transform
  ChiSquared(nu=1)           hi_model.py:95               fwhm2_norm_log__      x = exp(u),        log|J| = u
  Normal(0, 1)               hi_model.py:99-129           (untransformed)
  TruncatedNormal(mu, sigma, lower=0)  hi_physical_model.py:138   nth_fwhm_1pc_interval__   x = exp(u) + 0, log|J| = u
  HalfNormal(1)              e.g. emission_model.py:62    *_log__               x = exp(u),        log|J| = u
  Beta(a, b)                 e.g. emission_model.py:76    *_logodds__           x = sigmoid(u),    log|J| = log x + log(1-x)
  Uniform(0, 1)              emission_model.py:107        *_interval__          x = sigmoid(u),    log|J| = log x + log(1-x)
  Normal(mu(theta), sigma)   emission_absorption_model.py:53   (untransformed; mu depends on tkin / fwhm2_thermal)
  Normal(0, prior_baseline_coeffs[i])  bayes_spec add_baseline_priors (third-party)

The physical models' ``fwhm_L_norm`` is ONE HalfNormal variable per evaluation
(hi_physical_model.py:148), so its prior and Jacobian are counted once, not per cloud.
"""

import math

import numpy as np

from . import models_ref as mr
from .backend import backend_of

LOG_2PI = math.log(2.0 * math.pi)

# free RV -> (distribution, transform)
_DIST = {
    "fwhm2_norm": ("chi2_1", "log"),
    "velocity_norm": ("normal01", None),
    "log10_nHI_norm": ("normal01", None),
    "log10_n_alpha_norm": ("normal01", None),
    "nth_fwhm_1pc": ("truncnormal_lower0", "log"),
    "fwhm_L_norm": ("halfnormal1", "log"),
    "tau_total_norm": ("halfnormal1", "log"),
    "TB_fwhm_norm": ("halfnormal1", "log"),
    "NHI_fwhm2_thermal_norm": ("halfnormal1", "log"),
    "ff_NHI_norm": ("halfnormal1", "log"),
    "tkin_factor": ("beta", "logodds"),
    "tkin_factor_norm": ("beta", "logodds"),
    "fwhm2_thermal_fraction": ("beta", "logodds"),
    "fwhm2_thermal_fraction_norm": ("beta", "logodds"),
    "filling_factor_norm": ("uniform01", "logodds"),
    "log10_wt_ff_tkin": ("normal_coupled", None),
    "log10_wt_ff_fwhm2_thermal": ("normal_coupled", None),
}

PRIOR_DEFAULTS = {
    "prior_tkin_factor": [2.0, 2.0],  # absorption_model.py:25, emission_model.py:41
    "prior_fwhm2_thermal_fraction": [2.0, 2.0],  # absorption_physical_model.py:25
    "prior_baseline_coeffs": None,  # -> [1.0]*(baseline_degree+1) per dataset
}


def value_name(name):
    """PyMC's name of the unconstrained value variable of free RV ``name``."""
    dist, tr = _DIST[name]
    if tr is None:
        return name
    if dist == "truncnormal_lower0" or dist == "uniform01":
        return name + "_interval__"
    return name + ("_log__" if tr == "log" else "_logodds__")


def backward(name, u):
    """Unconstrained -> constrained value."""
    xp = backend_of(u)
    tr = _DIST[name][1]
    if tr is None:
        return u
    if tr == "log":
        return xp.exp(u)
    return 1.0 / (1.0 + xp.exp(-u))


def forward(name, x):
    """Constrained -> unconstrained value (NumPy; for building test inputs)."""
    tr = _DIST[name][1]
    if tr is None:
        return np.asarray(x, dtype=np.float64)
    if tr == "log":
        return np.log(x)
    return np.log(x) - np.log1p(-x)


def log_jacobian(name, u, x):
    xp = backend_of(u)
    tr = _DIST[name][1]
    if tr is None:
        return 0.0 * u
    if tr == "log":
        return u
    return xp.log(x) + xp.log(1.0 - x)


def _beta_params(spec, name):
    key = "prior_tkin_factor" if name.startswith("tkin_factor") else "prior_fwhm2_thermal_fraction"
    return spec.get(key, PRIOR_DEFAULTS[key])


def density(spec, name, x, det):
    """log density of free RV ``name`` at constrained value ``x`` (elementwise)."""
    xp = backend_of(x)
    dist = _DIST[name][0]
    if dist == "normal01":
        return -0.5 * x**2.0 - 0.5 * LOG_2PI
    if dist == "chi2_1":  # Gamma(alpha=1/2, beta=1/2)
        return -0.5 * xp.log(x) - 0.5 * x - 0.5 * math.log(2.0) - math.lgamma(0.5)
    if dist == "halfnormal1":
        return -0.5 * x**2.0 + 0.5 * math.log(2.0 / math.pi)
    if dist == "beta":
        a, b = _beta_params(spec, name)
        lnB = math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)
        return (a - 1.0) * xp.log(x) + (b - 1.0) * xp.log(1.0 - x) - lnB
    if dist == "uniform01":
        return 0.0 * x
    if dist == "truncnormal_lower0":
        mu, sigma = spec["prior_nth_fwhm_1pc"]
        alpha = (0.0 - mu) / sigma
        lccdf = math.log(0.5 * math.erfc(alpha / math.sqrt(2.0)))  # log(1 - Phi(alpha))
        return -0.5 * ((x - mu) / sigma) ** 2.0 - math.log(sigma) - 0.5 * LOG_2PI - lccdf
    if dist == "normal_coupled":
        sigma = spec["prior_sigma_log10_NHI"]
        T = det["tkin"] if name == "log10_wt_ff_tkin" else det["fwhm2_thermal"]
        mu = -math.log(10.0) * sigma**2.0 / 2.0 - xp.log10(T)
        return -0.5 * ((x - mu) / sigma) ** 2.0 - math.log(sigma) - 0.5 * LOG_2PI
    raise KeyError(dist)


def baseline_sigmas(spec, key, n_terms):
    p = spec.get("prior_baseline_coeffs")
    if p is None:
        return np.ones(n_terms)
    return np.asarray(p[key], dtype=np.float64)


def model_logp(spec, data, unconstrained, baseline):
    """Total model logp ``[...]`` in the unconstrained space: priors + Jacobians +
    likelihood.  ``unconstrained`` maps free-RV names to unconstrained values (``[..., N]``,
    ``[..., 1]`` for the physical models' fwhm_L_norm); ``baseline`` maps
    ``baseline_{key}_norm`` to ``[..., D+1]`` (Normal priors, untransformed)."""
    xp = backend_of(*unconstrained.values())
    free = {k: backward(k, u) for k, u in unconstrained.items()}
    det = mr.cloud_inputs(spec, free)
    total = mr.logp_from_inputs(spec, data, det, baseline)
    for name, u in unconstrained.items():
        x = free[name]
        term = density(spec, name, x, det) + log_jacobian(name, u, x)
        total = total + xp.sum(term, axis=-1)
    for key, coeff in baseline.items():
        dkey = key[len("baseline_"):-len("_norm")]
        sig = xp.asarray(baseline_sigmas(spec, dkey, coeff.shape[-1]))
        term = -0.5 * (coeff / sig) ** 2.0 - xp.log(sig) - 0.5 * LOG_2PI
        total = total + xp.sum(term, axis=-1)
    return total
