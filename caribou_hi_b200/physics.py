"""Drop-in for ``caribou_hi.physics``: same function names, argument order and meaning,
with the arithmetic executed by the CUDA library.

* With NumPy arrays / Python floats the functions that return PyTensor variables in the
  reference (``gaussian``, ``calc_spin_temp``, ``calc_log10_column_density``,
  ``calc_log10_nonthermal_pressure``, ``calc_pseudo_voigt``, ``radiative_transfer`` --
  the ones reference ``tests/test_physics.py`` calls ``.eval()`` on) return a
  :class:`DeviceValue` whose ``.eval()`` runs the kernel on ``cuda:0`` and returns NumPy.
* The helpers the reference writes in plain Python arithmetic (``lorentzian``,
  ``calc_thermal_fwhm2``, ``calc_kinetic_temp``, ``calc_nonthermal_fwhm``,
  ``calc_depth_nonthermal``, ``calc_log10_density``, ``calc_tau_total``) stay plain
  arithmetic on their arguments, exactly as in the reference (physics.py:109-188,210-244)
  -- they work unchanged on floats, NumPy arrays and PyTensor variables.
* With PyTensor variables as inputs (only possible where PyTensor is installed) the
  symbolic Ops of :mod:`caribou_hi_b200.ops` are returned instead.
"""

import numpy as np

from . import ops

_engine = None


def _default_engine():
    """Process-wide physics-level handle on cuda:0 (created on first use)."""
    global _engine
    if _engine is None:
        from .engine import Engine

        _engine = Engine("physics", 1)
    return _engine


class DeviceValue:
    """Deferred device evaluation with the ``.eval()`` protocol of a PyTensor variable."""

    def __init__(self, fn, *args):
        self._fn = fn
        self._args = args

    def eval(self, *_, **__):
        out = self._fn(*self._args)
        return out if out.ndim else out[()]


def _symbolic(*xs):
    return ops.HAVE_PYTENSOR and any(ops.is_variable(x) for x in xs)


def gaussian(x, center, fwhm):
    """Normalised Gaussian (physics.py:16-35)."""
    if _symbolic(x, center, fwhm):
        return ops.gaussian_graph(x, center, fwhm)
    return DeviceValue(lambda: _default_engine().elementwise("gaussian", x, center, fwhm))


def lorentzian(x, center, fwhm):
    """Normalised Lorentzian (physics.py:38-55): plain arithmetic, as in the reference."""
    return fwhm / (2.0 * np.pi) / ((x - center) ** 2.0 + (fwhm / 2.0) ** 2.0)


def calc_spin_temp(kinetic_temp, density, n_alpha):
    """Spin temperature, Kim et al. (2014) eq. 4 (physics.py:58-106)."""
    if _symbolic(kinetic_temp, density, n_alpha):
        return ops.SpinTempOp()(kinetic_temp, density, n_alpha)
    return DeviceValue(lambda: _default_engine().spin_temp(kinetic_temp, density, n_alpha))


def calc_thermal_fwhm2(kinetic_temp):
    """physics.py:109-124."""
    const = 0.2139  # km/s K-1/2
    return const**2 * kinetic_temp


def calc_kinetic_temp(thermal_fwhm2):
    """physics.py:127-143."""
    const = 0.2139  # km/s K-1/2
    return thermal_fwhm2 / const**2


def calc_nonthermal_fwhm(depth, nth_fwhm_1pc, depth_nth_fwhm_power):
    """physics.py:146-165."""
    return nth_fwhm_1pc * depth**depth_nth_fwhm_power


def calc_depth_nonthermal(fwhm_nonthermal, nth_fwhm_1pc, depth_nth_fwhm_power):
    """physics.py:168-188."""
    return (fwhm_nonthermal / nth_fwhm_1pc) ** (1.0 / depth_nth_fwhm_power)


def calc_log10_column_density(tau_total, spin_temp):
    """physics.py:191-207."""
    if _symbolic(tau_total, spin_temp):
        return ops.log10_column_density_graph(tau_total, spin_temp)
    return DeviceValue(lambda: _default_engine().elementwise("log10_column_density", tau_total, spin_temp))


def calc_log10_density(log10_column_density, log10_depth):
    """physics.py:210-225."""
    return log10_column_density - log10_depth - 18.489351


def calc_tau_total(column_density, spin_temp):
    """physics.py:228-244."""
    const = 1.82243e18  # cm-2 (K km s-1)-1
    return column_density / spin_temp / const


def calc_log10_nonthermal_pressure(log10_density, fwhm2_nonthermal):
    """physics.py:247-262."""
    if _symbolic(log10_density, fwhm2_nonthermal):
        return ops.log10_nonthermal_pressure_graph(log10_density, fwhm2_nonthermal)
    return DeviceValue(
        lambda: _default_engine().elementwise("log10_nonthermal_pressure", log10_density, fwhm2_nonthermal)
    )


def calc_pseudo_voigt(velo_axis, velocity, fwhm2, fwhm_L):
    """Pseudo-Voigt line profile, shape ``[S, N]`` (physics.py:265-309)."""
    if _symbolic(velocity, fwhm2, fwhm_L):
        return ops.PseudoVoigtOp(np.asarray(velo_axis, dtype=np.float64))(velocity, fwhm2, fwhm_L)
    return DeviceValue(lambda: _default_engine().pseudo_voigt(velo_axis, velocity, fwhm2, fwhm_L))


def radiative_transfer(tau, tspin, filling_factor, bg_temp):
    """Ordered-cloud radiative transfer, ON - OFF, shape ``[S]`` (physics.py:312-365)."""
    if _symbolic(tau, tspin, filling_factor):
        return ops.RadiativeTransferOp(float(bg_temp))(tau, tspin, filling_factor)
    return DeviceValue(lambda: _default_engine().radiative_transfer(tau, tspin, filling_factor, bg_temp))
