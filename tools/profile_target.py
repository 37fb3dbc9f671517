"""Small fixed workload for ncu captures: BASELINE configs[1] shape at a reduced number of
draws (20k), 3 logp+grad batches.  usage: python tools/profile_target.py [draws] [pv]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caribou_hi_b200 import Engine, synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
pv = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
kind, N, S = "emission_absorption_physical", 4, 1000
eng = Engine(kind, N, lorentzian=pv, prior_fwhm_L=1.0 if pv else 0.0, prior_amplitude=1e21)
v = np.linspace(-100, 100, S)
rng = np.random.default_rng(0)
eng.set_spectra(emission=dict(spectral=v, brightness=rng.normal(10, 5, S), noise=0.1, basis=np.full((1, S), 0.1)),
                absorption=dict(spectral=v, brightness=rng.normal(0.3, 0.1, S), noise=0.01, basis=np.full((1, S), 0.01)))
th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal"]
theta = eng.pack(synthetic.draw_free(kind, N, B, rng, lorentzian=pv, thermal_fn=th))
# device-resident call: ONE launch of each kernel per batch (the host entry point would cut the
# batch into pipeline pieces)
import torch
t_theta = torch.from_numpy(theta).cuda()
lp = torch.empty(B, dtype=torch.float64, device="cuda"); gr = torch.empty((B, 10, N), dtype=torch.float64, device="cuda")
ce = torch.zeros((B, 1), dtype=torch.float64, device="cuda"); ca = torch.zeros((B, 1), dtype=torch.float64, device="cuda")
ge = torch.empty_like(ce); ga = torch.empty_like(ca)
for _ in range(3):
    eng.logp_grad_dev(B, t_theta, lp, gr, coeff_e=ce, coeff_a=ca, gcoeff_e=ge, gcoeff_a=ga)
eng.synchronize()
print("ok", float(torch.isfinite(lp).double().mean()), eng.stage_ms(1))
