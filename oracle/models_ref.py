"""Oracle restatement of the six model classes' parameter maps and likelihoods
(test infrastructure; see ``oracle/__init__.py`` for the pinning status).

A model is described by a plain ``dict`` *spec*::

    {"kind": "emission_absorption_physical", "n_clouds": 4,
     "bg_temp": 3.77, "depth_nth_fwhm_power": 1/3,
     "prior_fwhm2": 500.0, "prior_velocity": [0, 10], ...,          # add_priors kwargs
     "prior_fwhm_L": None | float, "prior_sigma_log10_NHI": None | float}

and evaluated on a ``dict`` of *free* random-variable values named exactly as the
reference names its PyMC RVs (``fwhm2_norm``, ``velocity_norm`` ...), each of shape
``[..., N]`` (``fwhm_L_norm`` of the physical models is ``[..., 1]``: one value
per evaluation shared by all clouds, hi_physical_model.py:148-149).

``cloud_inputs`` returns the reference's Deterministics; ``predict`` the two
spectra; ``logp`` the summed ``pm.Normal`` log-likelihood.
"""

import numpy as np

from . import physics_ref as ph
from .backend import backend_of

KINDS = (
    "absorption",
    "emission",
    "emission_absorption",
    "absorption_physical",
    "emission_physical",
    "emission_absorption_physical",
)

# add_priors defaults: hi_model.py:57-65, hi_physical_model.py:70-78,
# absorption_model.py:22-28, emission_model.py:38-44, emission_absorption_model.py:23-28,
# absorption_physical_model.py:22-28, emission_physical_model.py:38-44,
# emission_absorption_physical_model.py:22-27; ctor defaults emission_model.py:24,
# hi_physical_model.py:24
DEFAULTS = {
    "prior_fwhm2": 500.0,
    "prior_log10_nHI": [0.0, 1.5],
    "prior_velocity": [0.0, 10.0],
    "prior_log10_n_alpha": [-6.0, 2.0],
    "prior_nth_fwhm_1pc": [1.75, 0.25],
    "prior_fwhm_L": None,
    "prior_tau_total": 10.0,
    "prior_TB_fwhm": 1000.0,
    "prior_ff_NHI": 1.0e21,
    "prior_NHI_fwhm2_thermal": 1.0e20,
    "prior_tkin_factor": [2.0, 2.0],
    "prior_fwhm2_thermal_fraction": [2.0, 2.0],
    "prior_sigma_log10_NHI": None,  # physical EA model default is 0.5 (set in make_spec)
    "bg_temp": 3.77,
    "depth_nth_fwhm_power": 1.0 / 3.0,
}


def make_spec(kind, n_clouds, **kwargs):
    assert kind in KINDS, kind
    spec = dict(DEFAULTS)
    if kind == "emission_absorption_physical":
        spec["prior_sigma_log10_NHI"] = 0.5  # emission_absorption_physical_model.py:24
    spec.update(kwargs)
    spec["kind"] = kind
    spec["n_clouds"] = int(n_clouds)
    return spec


def is_physical(spec):
    return spec["kind"].endswith("_physical")


def has_emission(spec):
    return spec["kind"].startswith("emission")


def has_absorption(spec):
    return "absorption" in spec["kind"]


def free_names(spec):
    """Names of the per-cloud free RVs of the model, in the reference's creation
    order (baseline coefficients excluded)."""
    k = spec["kind"]
    if is_physical(spec):
        names = ["fwhm2_norm", "velocity_norm", "log10_n_alpha_norm", "nth_fwhm_1pc"]
    else:
        names = ["fwhm2_norm", "log10_nHI_norm", "velocity_norm", "log10_n_alpha_norm"]
    if spec["prior_fwhm_L"] is not None:
        names.append("fwhm_L_norm")
    if k == "absorption":
        names += ["tau_total_norm", "tkin_factor"]
    elif k in ("emission", "emission_absorption"):
        names += ["TB_fwhm_norm", "tkin_factor_norm", "filling_factor_norm"]
        if k == "emission_absorption" and spec["prior_sigma_log10_NHI"] is not None:
            names.append("log10_wt_ff_tkin")
    elif k == "absorption_physical":
        names += ["NHI_fwhm2_thermal_norm", "fwhm2_thermal_fraction"]
    else:
        names += ["ff_NHI_norm", "fwhm2_thermal_fraction_norm", "filling_factor_norm"]
        if k == "emission_absorption_physical":
            names.append("log10_wt_ff_fwhm2_thermal")
    return names


def _common(spec, free, out):
    """hi_model.py:93-136 / hi_physical_model.py:109-151 (shared deterministics)."""
    xp = backend_of(*free.values())
    out["fwhm2"] = spec["prior_fwhm2"] * free["fwhm2_norm"]
    out["velocity"] = spec["prior_velocity"][0] + spec["prior_velocity"][1] * free["velocity_norm"]
    out["log10_n_alpha"] = (
        spec["prior_log10_n_alpha"][0] + spec["prior_log10_n_alpha"][1] * free["log10_n_alpha_norm"]
    )
    if not is_physical(spec):
        out["log10_nHI"] = (
            spec["prior_log10_nHI"][0] + spec["prior_log10_nHI"][1] * free["log10_nHI_norm"]
        )
    else:
        out["nth_fwhm_1pc"] = free["nth_fwhm_1pc"]
    if spec["prior_fwhm_L"] is not None:
        out["fwhm_L"] = spec["prior_fwhm_L"] * free["fwhm_L_norm"]
    else:
        out["fwhm_L"] = xp.zeros_like(out["fwhm2"][..., :1])  # pm.Data("fwhm_L", 0.0)
    return xp


def _physical_deterministics(spec, out, xp):
    """hi_physical_model.py:153-220."""
    out["fwhm2_nonthermal"] = out["fwhm2"] - out["fwhm2_thermal"]
    out["depth"] = ph.calc_depth_nonthermal(
        xp.sqrt(out["fwhm2_nonthermal"]), out["nth_fwhm_1pc"], spec["depth_nth_fwhm_power"]
    )
    out["log10_nHI"] = ph.calc_log10_density(out["log10_NHI"], xp.log10(out["depth"]))
    out["tspin"] = ph.calc_spin_temp(
        out["tkin"], 10.0 ** out["log10_nHI"], 10.0 ** out["log10_n_alpha"]
    )
    out["tau_total"] = ph.calc_tau_total(10.0 ** out["log10_NHI"], out["tspin"])
    out["log10_Pth"] = xp.log10(out["tkin"]) + out["log10_nHI"]
    out["log10_Pnth"] = ph.calc_log10_nonthermal_pressure(out["log10_nHI"], out["fwhm2_nonthermal"])
    out["log10_Ptot"] = xp.log10(10.0 ** out["log10_Pth"] + 10.0 ** out["log10_Pnth"])


def _observational_deterministics(out, xp):
    """hi_model.py:138-162."""
    out["log10_Pth"] = xp.log10(out["tkin"]) + out["log10_nHI"]
    out["log10_NHI"] = ph.calc_log10_column_density(out["tau_total"], out["tspin"])
    out["log10_depth"] = out["log10_NHI"] - out["log10_nHI"] - ph.PC_CM_LOG10


def cloud_inputs(spec, free):
    """Free RV values -> every reference Deterministic (dict of ``[..., N]``)."""
    out = {}
    xp = _common(spec, free, out)
    k = spec["kind"]
    const_G = np.sqrt(2.0 * np.pi) / (2.0 * np.sqrt(2.0 * np.log(2.0)))

    if k == "absorption":
        # absorption_model.py:44-72
        out["tau_total"] = spec["prior_tau_total"] * free["tau_total_norm"]
        tkin_max = ph.calc_kinetic_temp(out["fwhm2"])
        out["tkin"] = free["tkin_factor"] * tkin_max
        out["tspin"] = ph.calc_spin_temp(
            out["tkin"], 10.0 ** out["log10_nHI"], 10.0 ** out["log10_n_alpha"]
        )
        _observational_deterministics(out, xp)

    elif k in ("emission", "emission_absorption"):
        # emission_model.py:60-132
        out["TB_fwhm"] = spec["prior_TB_fwhm"] * free["TB_fwhm_norm"]
        tkin_min = out["TB_fwhm"] / xp.sqrt(out["fwhm2"])
        tkin_max = ph.calc_kinetic_temp(out["fwhm2"])
        out["tkin"] = xp.where(
            tkin_min > tkin_max,
            tkin_max,
            free["tkin_factor_norm"] * (tkin_max - tkin_min) + tkin_min,
        )
        out["tspin"] = ph.calc_spin_temp(
            out["tkin"], 10.0 ** out["log10_nHI"], 10.0 ** out["log10_n_alpha"]
        )
        ff_min = tkin_min / out["tspin"]
        out["filling_factor"] = xp.where(
            ff_min > 1.0,
            xp.ones_like(ff_min),
            free["filling_factor_norm"] * (1.0 - ff_min) + ff_min,
        )
        exp_tau_peak = ff_min / out["filling_factor"]
        exp_tau_peak = xp.clip(exp_tau_peak, 0.0, 0.9999)
        tau_peak = -xp.log(1.0 - exp_tau_peak)
        out["tau_total"] = tau_peak * xp.sqrt(out["fwhm2"]) * const_G
        if k == "emission_absorption":
            # emission_absorption_model.py:45-64
            if spec["prior_sigma_log10_NHI"] is None:
                out["absorption_weight"] = xp.ones_like(out["fwhm2"])
            else:
                out["absorption_weight"] = (
                    10.0 ** free["log10_wt_ff_tkin"] * out["filling_factor"] * out["tkin"]
                )
        _observational_deterministics(out, xp)

    elif k == "absorption_physical":
        # absorption_physical_model.py:43-77
        NHI_fwhm2_thermal = spec["prior_NHI_fwhm2_thermal"] * free["NHI_fwhm2_thermal_norm"]
        out["fwhm2_thermal"] = out["fwhm2"] * free["fwhm2_thermal_fraction"]
        out["tkin"] = ph.calc_kinetic_temp(out["fwhm2_thermal"])
        out["log10_NHI"] = xp.log10(NHI_fwhm2_thermal * out["fwhm2_thermal"])
        _physical_deterministics(spec, out, xp)

    else:
        # emission_physical_model.py:59-129
        ff_NHI = spec["prior_ff_NHI"] * free["ff_NHI_norm"]
        tkin_min = (ff_NHI / ph.COLUMN_CONST) / (const_G * xp.sqrt(out["fwhm2"]))
        frac_min = ph.calc_thermal_fwhm2(tkin_min) / out["fwhm2"]
        out["fwhm2_thermal_fraction"] = xp.where(
            frac_min > 1.0,
            xp.ones_like(frac_min),
            free["fwhm2_thermal_fraction_norm"] * (1.0 - frac_min) + frac_min,
        )
        out["fwhm2_thermal"] = out["fwhm2"] * out["fwhm2_thermal_fraction"]
        out["tkin"] = ph.calc_kinetic_temp(out["fwhm2_thermal"])
        ff_min = tkin_min / out["tkin"]
        out["filling_factor"] = xp.where(
            ff_min > 1.0,
            xp.ones_like(ff_min),
            free["filling_factor_norm"] * (1.0 - ff_min) + ff_min,
        )
        out["log10_NHI"] = xp.log10(ff_NHI / out["filling_factor"])
        _physical_deterministics(spec, out, xp)
        if k == "emission_absorption_physical":
            # emission_absorption_physical_model.py:41-60
            out["absorption_weight"] = (
                10.0 ** free["log10_wt_ff_fwhm2_thermal"]
                * out["filling_factor"]
                * out["fwhm2_thermal"]
            )
    return out


def predict_from_inputs(spec, data, det, baseline=None):
    """The ``add_likelihood`` bodies: absorption_model.py:80-96, emission_model.py:
    140-160, emission_absorption_model.py:69-105 and the physical equivalents
    (absorption_physical_model.py:85-101, emission_physical_model.py:137-157,
    emission_absorption_physical_model.py:65-101).

    ``data`` maps "emission"/"absorption" to ``OracleData``; ``baseline`` maps
    ``baseline_{key}_norm`` to coefficient arrays ``[..., D+1]`` (zeros if None).
    Returns a dict of predicted spectra ``[..., S]``.
    """
    out = {}
    if has_emission(spec):
        d = data["emission"]
        prof = ph.calc_pseudo_voigt(d.spectral, det["velocity"], det["fwhm2"], det["fwhm_L"])
        tau = det["tau_total"][..., None, :] * prof
        pred = ph.radiative_transfer(tau, det["tspin"], det["filling_factor"], spec["bg_temp"])
        out["emission"] = pred + d.predict_baseline(_coeffs(baseline, "emission", pred))
    if has_absorption(spec):
        d = data["absorption"]
        prof = ph.calc_pseudo_voigt(d.spectral, det["velocity"], det["fwhm2"], det["fwhm_L"])
        if has_emission(spec):
            tau = det["absorption_weight"][..., None, :] * det["tau_total"][..., None, :] * prof
        else:
            tau = det["tau_total"][..., None, :] * prof
        pred = ph.absorption_spectrum(tau)
        out["absorption"] = pred + d.predict_baseline(_coeffs(baseline, "absorption", pred))
    return out


def _coeffs(baseline, key, like):
    if baseline is None:
        return None
    return baseline.get(f"baseline_{key}_norm")


def predict(spec, data, free, baseline=None):
    return predict_from_inputs(spec, data, cloud_inputs(spec, free), baseline)


def normal_logp(value, mu, sigma):
    """PyMC ``Normal.logp`` (third-party, pymc>=5; restated from its published
    definition): ``-0.5*((value-mu)/sigma)**2 - log(sqrt(2*pi)) - log(sigma)``,
    summed over channels by the model logp."""
    xp = backend_of(mu)
    sigma = xp.asarray(sigma)
    value = xp.asarray(value)
    return -0.5 * ((value - mu) / sigma) ** 2.0 - np.log(np.sqrt(2.0 * np.pi)) - xp.log(sigma)


def logp_from_spectra(data, spectra):
    total = 0.0
    for key, mu in spectra.items():
        d = data[key]
        xp = backend_of(mu)
        total = total + xp.sum(normal_logp(d.brightness, mu, d.noise), axis=-1)
    return total


def logp(spec, data, free, baseline=None):
    """Summed likelihood logp of every observed dataset, shape ``[...]``."""
    return logp_from_spectra(data, predict(spec, data, free, baseline))


def logp_from_inputs(spec, data, det, baseline=None):
    return logp_from_spectra(data, predict_from_inputs(spec, data, det, baseline))


class OracleData:
    """Minimal stand-in for ``bayes_spec.SpecData`` (third-party, bayes_spec>=1.7.8;
    UNVERIFIED restatement from its published source): keeps the spectral axis, the
    brightness and the (scalar or per-channel) noise, the normalised spectral axis
    ``(x-mean)/std`` and the brightness un-normalisation ``x*median(noise) +
    median(brightness)`` used by ``BaseModel.predict_baseline``."""

    def __init__(self, spectral, brightness, noise):
        self.spectral = np.asarray(spectral, dtype=np.float64)
        self.brightness = np.asarray(brightness, dtype=np.float64)
        if np.ndim(noise) == 0:
            noise = np.ones_like(self.brightness) * float(noise)
        self.noise = np.asarray(noise, dtype=np.float64)
        self._spectral_offset = np.mean(self.spectral)
        self._spectral_scale = np.std(self.spectral)
        self.spectral_norm = (self.spectral - self._spectral_offset) / self._spectral_scale
        self._brightness_offset = np.median(self.brightness)
        self._brightness_scale = np.median(self.noise)

    def predict_baseline(self, coeffs):
        """``sum_i coeff_i/(i+1)**i * spectral_norm**i`` un-normalised to brightness
        units (bayes_spec ``BaseModel.predict_baseline``; UNVERIFIED third-party)."""
        if coeffs is None:
            return self._brightness_offset
        xp = backend_of(coeffs)
        x = xp.asarray(self.spectral_norm)
        n_terms = coeffs.shape[-1]
        terms = [coeffs[..., i : i + 1] / (i + 1.0) ** i * x**i for i in range(n_terms)]
        total = terms[0]
        for t in terms[1:]:
            total = total + t
        return total * self._brightness_scale + self._brightness_offset
