"""Parameter-block layout shared by the host code and ``libchi`` (``include/chi.h``).

``theta[B, CHI_NSLOT, N]`` holds, per draw and cloud, either the seven likelihood
inputs (``MODEL_PHYSICS``) or the free random variables of a reference model class,
named exactly as the reference names its PyMC RVs.
"""

from . import _lib

NSLOT = _lib.CHI_NSLOT

KINDS = {
    "physics": _lib.MODEL_PHYSICS,
    "absorption": _lib.MODEL_ABSORPTION,
    "emission": _lib.MODEL_EMISSION,
    "emission_absorption": _lib.MODEL_EMISSION_ABSORPTION,
    "absorption_physical": _lib.MODEL_ABSORPTION_PHYSICAL,
    "emission_physical": _lib.MODEL_EMISSION_PHYSICAL,
    "emission_absorption_physical": _lib.MODEL_EMISSION_ABSORPTION_PHYSICAL,
}

PHYSICS_SLOTS = {
    "velocity": 0,
    "fwhm2": 1,
    "fwhm_L": 2,
    "tau_total": 3,
    "tspin": 4,
    "filling_factor": 5,
    "absorption_weight": 6,
}

# slot -> RV name, per model kind (reference file:line in include/chi.h)
_AMPLITUDE = {
    "absorption": "tau_total_norm",
    "emission": "TB_fwhm_norm",
    "emission_absorption": "TB_fwhm_norm",
    "absorption_physical": "NHI_fwhm2_thermal_norm",
    "emission_physical": "ff_NHI_norm",
    "emission_absorption_physical": "ff_NHI_norm",
}
_THERMAL = {
    "absorption": "tkin_factor",
    "emission": "tkin_factor_norm",
    "emission_absorption": "tkin_factor_norm",
    "absorption_physical": "fwhm2_thermal_fraction",
    "emission_physical": "fwhm2_thermal_fraction_norm",
    "emission_absorption_physical": "fwhm2_thermal_fraction_norm",
}
_WEIGHT = {
    "emission_absorption": "log10_wt_ff_tkin",
    "emission_absorption_physical": "log10_wt_ff_fwhm2_thermal",
}

DET_NAMES = [
    "velocity",
    "fwhm2",
    "fwhm_L",
    "tau_total",
    "tspin",
    "filling_factor",
    "absorption_weight",
    "tkin",
    "log10_NHI",
    "log10_nHI",
    "log10_depth",
    "fwhm2_thermal",
    "fwhm2_nonthermal",
    "log10_Pth",
    "log10_Pnth",
    "log10_Ptot",
    "log10_n_alpha",
    "depth",
    "TB_fwhm",
    "fwhm2_thermal_fraction",
]


def is_physical(kind):
    return kind.endswith("_physical")


def has_emission(kind):
    return kind.startswith("emission")


def has_absorption(kind):
    return "absorption" in kind


def slot_names(kind, lorentzian, free_weight=True):
    """``{rv_name: slot}`` of the free RVs the model kind actually has."""
    if kind == "physics":
        return dict(PHYSICS_SLOTS)
    s = {"fwhm2_norm": 0, "velocity_norm": 1, "log10_n_alpha_norm": 3}
    if is_physical(kind):
        s["nth_fwhm_1pc"] = 4
    else:
        s["log10_nHI_norm"] = 2
    if lorentzian:
        s["fwhm_L_norm"] = 5
    s[_AMPLITUDE[kind]] = 6
    s[_THERMAL[kind]] = 7
    if has_emission(kind):
        s["filling_factor_norm"] = 8
    if kind in _WEIGHT and (free_weight or kind == "emission_absorption_physical"):
        s[_WEIGHT[kind]] = 9
    return s


_LOG = {"fwhm2_norm", "fwhm_L_norm", "tau_total_norm", "TB_fwhm_norm", "NHI_fwhm2_thermal_norm", "ff_NHI_norm"}
_LOGODDS = {"tkin_factor", "tkin_factor_norm", "fwhm2_thermal_fraction", "fwhm2_thermal_fraction_norm"}
_INTERVAL = {"nth_fwhm_1pc", "filling_factor_norm"}


def value_name(rv):
    """PyMC's name for the unconstrained value variable of free RV ``rv`` (its default
    transform: log for positive, logodds for Beta, interval for bounded RVs)."""
    if rv in _LOG:
        return rv + "_log__"
    if rv in _LOGODDS:
        return rv + "_logodds__"
    if rv in _INTERVAL:
        return rv + "_interval__"
    return rv


def scalar_rvs(kind):
    """RVs that are one value per draw (shared by all clouds) in the reference:
    ``fwhm_L_norm`` of the physical models (hi_physical_model.py:148-149)."""
    return {"fwhm_L_norm"} if is_physical(kind) else set()
