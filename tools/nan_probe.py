import sys, numpy as np
sys.path.insert(0, "/root/repo")
from caribou_hi_b200 import Engine
z = np.load("/root/repo/tests/golden/ref_emission_absorption_physical_gauss.npz", allow_pickle=True)
b = 1
eng = Engine("physics", 3, lorentzian=True, has_emission=True, has_absorption=True)
eng.set_spectra(emission=dict(spectral=z["data_emission_spectral"], brightness=z["data_emission_brightness"], noise=z["data_emission_noise"]),
                absorption=dict(spectral=z["data_absorption_spectral"], brightness=z["data_absorption_brightness"], noise=z["data_absorption_noise"]))
names = ["velocity", "fwhm2", "fwhm_L", "tau_total", "tspin", "filling_factor", "absorption_weight"]
inp = {n: (z[f"det_{n}"][b][None, :] if f"det_{n}" in z.files else np.zeros((1, 3))) for n in names}
print({n: inp[n][0] for n in names})
mu_e, mu_a = eng.spectra(eng.pack(inp))
print("mu_e finite", np.isfinite(mu_e).sum(), "of", mu_e.size, "mu_a finite", np.isfinite(mu_a).sum(), "of", mu_a.size)
print("ref mu_e finite", np.isfinite(z["mu_emission"][b]).sum(), "mu_a", np.isfinite(z["mu_absorption"][b]).sum())
print(mu_a[0][:6])
