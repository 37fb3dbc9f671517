"""Oracle restatement of the reference's physics helpers (test infrastructure).

Every function follows the operation order of ``caribou_hi/physics.py`` in the
reference tree (cited per function) so that fp64 results agree with the PyTensor
graph to rounding.  Arguments may be NumPy arrays or torch fp64 tensors; the latter
give gradients through ``torch.autograd``.  All cloud-shaped arguments have the
clouds on the LAST axis and may carry leading batch axes; the spectral axis, where
present, is the second-to-last axis of the result (``[..., S, N]``).
"""

import numpy as np

from .backend import backend_of

LN2 = np.log(2.0)

# constants as written in the reference
T_RADIATION = 3.77  # physics.py:75
LYA_COEFF = 5.9e11  # physics.py:79
T_0 = 0.0681  # physics.py:82
A_10 = 2.8843e-15  # physics.py:85
THERMAL_CONST = 0.2139  # physics.py:123,142
COLUMN_CONST = 1.82243e18  # physics.py:206,243
PC_CM_LOG10 = 18.489351  # physics.py:225
PNTH_CONST = 671.8  # physics.py:262


def gaussian(x, center, fwhm):
    """physics.py:16-35 -- normalised Gaussian of the given FWHM."""
    xp = backend_of(x, center, fwhm)
    return xp.exp(-4.0 * LN2 * (x - center) ** 2.0 / fwhm**2.0) * xp.sqrt(
        4.0 * LN2 / (np.pi * fwhm**2.0)
    )


def lorentzian(x, center, fwhm):
    """physics.py:38-55 -- normalised Lorentzian of the given FWHM."""
    return fwhm / (2.0 * np.pi) / ((x - center) ** 2.0 + (fwhm / 2.0) ** 2.0)


def calc_spin_temp(kinetic_temp, density, n_alpha):
    """physics.py:58-106 -- Kim+14 eq. 4 with the Draine (2011) piecewise k10.

    Both branches of the ``switch`` are evaluated, the cold branch is taken for
    strictly ``kinetic_temp < 300``.
    """
    xp = backend_of(kinetic_temp, density, n_alpha)
    y_alpha = LYA_COEFF * n_alpha / kinetic_temp**1.5
    T2 = kinetic_temp / 100.0
    cold = 1.19e-10 * T2 ** (0.74 - 0.20 * xp.log(T2))
    warm = 2.24e-10 * T2**0.207 * xp.exp(-0.876 / T2)
    k10 = xp.where(kinetic_temp < 300.0, cold, warm)
    R10 = density * k10
    y_c = T_0 * R10 / (kinetic_temp * A_10)
    return (T_RADIATION + y_c * kinetic_temp + y_alpha * kinetic_temp) / (
        1 + y_c + y_alpha
    )


def calc_thermal_fwhm2(kinetic_temp):
    """physics.py:109-124."""
    return THERMAL_CONST**2 * kinetic_temp


def calc_kinetic_temp(thermal_fwhm2):
    """physics.py:127-143."""
    return thermal_fwhm2 / THERMAL_CONST**2


def calc_nonthermal_fwhm(depth, nth_fwhm_1pc, depth_nth_fwhm_power):
    """physics.py:146-165."""
    return nth_fwhm_1pc * depth**depth_nth_fwhm_power


def calc_depth_nonthermal(fwhm_nonthermal, nth_fwhm_1pc, depth_nth_fwhm_power):
    """physics.py:168-188."""
    return (fwhm_nonthermal / nth_fwhm_1pc) ** (1.0 / depth_nth_fwhm_power)


def calc_log10_column_density(tau_total, spin_temp):
    """physics.py:191-207."""
    xp = backend_of(tau_total, spin_temp)
    return xp.log10(xp.asarray(COLUMN_CONST * spin_temp * tau_total))


def calc_log10_density(log10_column_density, log10_depth):
    """physics.py:210-225."""
    return log10_column_density - log10_depth - PC_CM_LOG10


def calc_tau_total(column_density, spin_temp):
    """physics.py:228-244."""
    return column_density / spin_temp / COLUMN_CONST


def calc_log10_nonthermal_pressure(log10_density, fwhm2_nonthermal):
    """physics.py:247-262."""
    xp = backend_of(log10_density, fwhm2_nonthermal)
    return np.log10(PNTH_CONST) + log10_density + xp.log10(xp.asarray(fwhm2_nonthermal))


def calc_pseudo_voigt(velo_axis, velocity, fwhm2, fwhm_L):
    """physics.py:265-309 -- pseudo-Voigt profile on ``velo_axis`` ([S]) for clouds
    ``[..., N]``; returns ``[..., S, N]``.  ``fwhm_L`` is a scalar, ``[..., 1]`` or
    ``[..., N]``.  Both the Lorentzian and Gaussian parts use the same convolved
    width, which includes the channel-width term taken from the first two channels.
    """
    xp = backend_of(velo_axis, velocity, fwhm2, fwhm_L)
    velo_axis = xp.asarray(velo_axis)
    channel_size = xp.abs(velo_axis[1] - velo_axis[0])
    channel_fwhm = 4.0 * LN2 * channel_size / np.pi
    fwhm_conv = xp.sqrt(fwhm2 + channel_fwhm**2.0 + fwhm_L**2.0)
    frac = fwhm_L / fwhm_conv
    eta = 1.36603 * frac - 0.47719 * frac**2.0 + 0.11116 * frac**3.0
    # [..., 1, N] against [S, 1]
    ctr = velocity[..., None, :]
    wid = fwhm_conv[..., None, :]
    eta = eta[..., None, :]
    x = velo_axis[:, None]
    gauss_part = gaussian(x, ctr, wid)
    lorentz_part = lorentzian(x, ctr, wid)
    return eta * lorentz_part + (1.0 - eta) * gauss_part


def radiative_transfer(tau, tspin, filling_factor, bg_temp):
    """physics.py:312-365 -- ON minus OFF brightness through N ordered clouds
    (cloud 0 nearest).  ``tau`` is ``[..., S, N]``, ``tspin`` ``[..., N]``,
    ``filling_factor`` ``[..., N]`` or a scalar."""
    xp = backend_of(tau, tspin, filling_factor)
    if getattr(filling_factor, "ndim", 0) > 0:
        filling_factor = filling_factor[..., None, :]
    tspin = xp.asarray(tspin)[..., None, :]
    attenuation = xp.concatenate(
        [
            xp.ones_like(tau[..., 0:1]),
            xp.cumprod(1.0 - filling_factor + filling_factor * xp.exp(-tau), axis=-1),
        ],
        axis=-1,
    )
    emission_bg_attenuated = bg_temp * attenuation[..., -1]
    emission_clouds = filling_factor * tspin * (1.0 - xp.exp(-tau))
    emission_clouds_attenuated = emission_clouds * attenuation[..., :-1]
    emission = emission_bg_attenuated + xp.sum(emission_clouds_attenuated, axis=-1)
    return emission - bg_temp


def absorption_spectrum(tau):
    """absorption_model.py:91, emission_absorption_model.py:99 (and the physical
    equivalents): ``1 - exp(-sum_n tau)`` for ``tau`` ``[..., S, N]``."""
    xp = backend_of(tau)
    return 1.0 - xp.exp(-xp.sum(tau, axis=-1))
