"""BASELINE configs[2] shape (64 chains x 8 clouds x 2048+2048 channels, pseudo-Voigt): a few
logp+grad evaluations, for launch lists / latency work.  usage: python tools/small_batch_target.py [reps] [small_path]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caribou_hi_b200 import Engine, synthetic

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
path = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # chi_set_option small_path: 0 automatic, 1 pipeline, 2 one launch
kind, N, S, B = "emission_absorption_physical", 8, 2048, 64
eng = Engine(kind, N, lorentzian=True, prior_fwhm_L=1.0, prior_amplitude=1e21)
eng.set_option("small_path", path)
v = np.linspace(-100, 100, S)
rng = np.random.default_rng(0)
eng.set_spectra(emission=dict(spectral=v, brightness=rng.normal(10, 5, S), noise=0.1, basis=np.full((1, S), 0.1)),
                absorption=dict(spectral=v, brightness=rng.normal(0.3, 0.1, S), noise=0.01, basis=np.full((1, S), 0.01)))
th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal"]
theta = torch.from_numpy(eng.pack(synthetic.draw_free(kind, N, B, rng, lorentzian=True, thermal_fn=th))).cuda()
lp = torch.empty(B, dtype=torch.float64, device="cuda"); gr = torch.empty((B, 10, N), dtype=torch.float64, device="cuda")
ce = torch.zeros((B, 1), dtype=torch.float64, device="cuda"); ca = torch.zeros((B, 1), dtype=torch.float64, device="cuda")
ge = torch.empty_like(ce); ga = torch.empty_like(ca)
stream = torch.cuda.ExternalStream(eng.stream)
for _ in range(3):
    eng.logp_grad_dev(B, theta, lp, gr, coeff_e=ce, coeff_a=ca, gcoeff_e=ge, gcoeff_a=ga)
eng.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps):
    eng.logp_grad_dev(B, theta, lp, gr, coeff_e=ce, coeff_a=ca, gcoeff_e=ge, gcoeff_a=ga)
e1.record(stream); eng.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
st = eng.stage_ms(min(reps, 32)).mean(axis=0) * 1e3
print(f"B={B} N={N} S={S}+{S} pv=1: {us:.1f} us per evaluation back to back ({B*N*2*S/us/1e3:.1f} Gunits/s); "
      f"stages map {st[0]:.1f} moments {st[1]:.1f} abs {st[2]:.1f} epilogue {st[3]:.1f} us; "
      f"checksum {float(torch.nan_to_num(lp).sum()):.6e} {float(torch.nan_to_num(gr).sum()):.6e}", flush=True)
