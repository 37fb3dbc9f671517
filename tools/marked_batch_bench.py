"""What a large batch costs when EVERY draw is marked for the careful copy of the channel kernel (a negative
optical depth in each draw: the worst case of the two-launch form, csrc/moments_ea.cuh MODE) against the
one-launch form and against an unmarked batch (development aid).  usage: python tools/marked_batch_bench.py [draws]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caribou_hi_b200 import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
N, S = 4, 1000
rng = np.random.default_rng(0)
v = np.linspace(-100, 100, S)
eng = Engine("physics", N, lorentzian=False)
eng.set_spectra(emission=dict(spectral=v, brightness=rng.normal(10, 5, S), noise=0.1),
                absorption=dict(spectral=v, brightness=rng.normal(0.3, 0.1, S), noise=0.01))
inp = {"velocity": rng.normal(0, 10, (B, N)), "fwhm2": rng.uniform(5, 500, (B, N)), "fwhm_L": np.zeros((B, N)),
       "tau_total": rng.uniform(0.05, 8, (B, N)), "tspin": rng.uniform(30, 3000, (B, N)),
       "filling_factor": rng.uniform(0.1, 1.0, (B, N)), "absorption_weight": rng.uniform(0.3, 1.2, (B, N))}
lp = torch.empty(B, dtype=torch.float64, device="cuda"); gr = torch.empty((B, 10, N), dtype=torch.float64, device="cuda")
for label, frac in (("no draw marked", 0.0), ("1 % of the draws marked", 0.01), ("every draw marked", 1.0)):
    x = {k: a.copy() for k, a in inp.items()}
    n_mark = int(frac * B)
    if n_mark:
        x["tau_total"][rng.choice(B, n_mark, replace=False), 0] = -0.01
    theta = torch.from_numpy(eng.pack(x)).cuda()
    for one_pass in (0, 1):
        eng.set_option("ea_one_pass", one_pass)
        for burst in range(2):  # (the library learns the density of marks from calls that have COMPLETED)
            for _ in range(6):
                eng.logp_grad_dev(B, theta, lp, gr)
            eng.synchronize()
        st = eng.stage_ms(4).mean(axis=0)
        print(f"{label:26s} {'one launch ' if one_pass else 'two launches'}: channel stage {st[1]:.3f} ms, step {st.sum():.3f} ms", flush=True)
eng.close()
