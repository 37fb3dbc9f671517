"""Times the posterior-predictive (spectra-only) path at BASELINE configs[4] shape per GPU."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caribou_hi_b200 import Engine, synthetic

def run(B, N, S, pv):
    kind = "emission_absorption_physical"
    eng = Engine(kind, N, lorentzian=pv, prior_fwhm_L=1.0 if pv else 0.0, prior_amplitude=1e21)
    v = np.linspace(-100, 100, S); rng = np.random.default_rng(0)
    eng.set_spectra(emission=dict(spectral=v, brightness=np.zeros(S), noise=0.1, basis=np.full((1, S), 0.1)),
                    absorption=dict(spectral=v, brightness=np.zeros(S), noise=0.01, basis=np.full((1, S), 0.01)))
    th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal"]
    theta = torch.from_numpy(eng.pack(synthetic.draw_free(kind, N, B, rng, lorentzian=pv, thermal_fn=th))).cuda()
    mu_e = torch.empty((B, S), dtype=torch.float64, device="cuda"); mu_a = torch.empty((B, S), dtype=torch.float64, device="cuda")
    stream = torch.cuda.ExternalStream(eng.stream)
    for _ in range(2): eng.spectra_dev(B, theta, mu_e, mu_a)
    eng.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5): eng.spectra_dev(B, theta, mu_e, mu_a)
    e1.record(stream); eng.synchronize()
    ms = e0.elapsed_time(e1) / 5
    units = B * N * 2 * S
    print(f"spectra B={B} N={N} S={S}+{S} pv={int(pv)}: {ms:.3f} ms -> {units/ms/1e6:.1f} Gunits/s, output {2*B*S*8/ms/1e6:.0f} GB/s", flush=True)
    eng.close()

if __name__ == "__main__":
    run(25000, 8, 4096, True)
    run(25000, 8, 4096, False)
    run(100000, 4, 1000, False)
