"""Host-side stand-in for ``bayes_spec.SpecData`` and the polynomial baseline of
``bayes_spec.BaseModel.predict_baseline`` (third-party, bayes_spec>=1.7.8, NOT in the
reference tree and not installable here -- restated from its published source and
therefore UNVERIFIED; when the real bayes_spec is present use
``baseline_from_bayes_spec`` in INTEGRATION.md, which extracts the same affine form
by evaluation instead of trusting this restatement).

The CUDA library only ever sees the affine form ``mu += offset[s] + sum_i coeff_i *
basis_i[s]`` (``include/chi.h: chi_dataset``), so a different normalisation changes
this file only.
"""

import numpy as np


class SpecData:
    """Spectral axis, brightness and (scalar or per-channel) noise of one dataset.

    Mirrors the attributes the reference touches: ``.spectral``, ``.brightness``,
    ``.noise`` (emission_model.py:141,167-168) plus the normalisations used by the
    baseline model."""

    def __init__(self, spectral, brightness, noise, xlabel="Spectral", ylabel="Brightness"):
        self.spectral = np.asarray(spectral, dtype=np.float64)
        self.brightness = np.asarray(brightness, dtype=np.float64)
        if self.spectral.ndim != 1 or self.spectral.shape != self.brightness.shape:
            raise ValueError("size mismatch between brightness and spectral")
        if np.ndim(noise) == 0:
            noise = np.ones_like(self.brightness) * float(noise)
        self.noise = np.asarray(noise, dtype=np.float64)
        if self.noise.shape != self.spectral.shape:
            raise ValueError("size mismatch between noise and spectral")
        self.xlabel = xlabel
        self.ylabel = ylabel
        self._spectral_offset = float(np.mean(self.spectral))
        self._spectral_scale = float(np.std(self.spectral))
        self._brightness_offset = float(np.median(self.brightness))
        self._brightness_scale = float(np.median(self.noise))
        self.spectral_norm = (self.spectral - self._spectral_offset) / self._spectral_scale

    def baseline_basis(self, baseline_degree):
        """``basis [D+1, S]`` and scalar ``offset`` such that the baseline is
        ``offset + sum_i coeff_i * basis_i``: term i of predict_baseline is
        ``coeff_i / (i+1)**i * spectral_norm**i`` un-normalised by
        ``x * median(noise) + median(brightness)``."""
        basis = np.stack(
            [self.spectral_norm**i / (i + 1.0) ** i * self._brightness_scale for i in range(baseline_degree + 1)]
        )
        return basis, self._brightness_offset

    def as_dataset(self, baseline_degree):
        basis, offset = self.baseline_basis(baseline_degree)
        return {
            "spectral": self.spectral,
            "brightness": self.brightness,
            "noise": self.noise,
            "basis": basis,
            "offset": offset,
        }
