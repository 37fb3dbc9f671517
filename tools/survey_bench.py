"""BASELINE configs[3] shape on one GPU: 10^4 sightlines x 5 clouds x 1000 channels, `per` chains
per sightline, device-resident inputs; prints the stage times.  usage: python tools/survey_bench.py [per]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caribou_hi_b200 import Engine, synthetic

per = int(sys.argv[1]) if len(sys.argv) > 1 else 2
kind, N, S, n_sl = "emission_absorption_physical", 5, 1000, 10_000
rng = np.random.default_rng(8)
v = np.linspace(-100.0, 100.0, S)
ye = rng.normal(10, 5, (n_sl, S)); ya = rng.normal(0.3, 0.1, (n_sl, S))
eng = Engine(kind, N, prior_amplitude=1e21)
eng.set_spectra(emission=dict(spectral=v, brightness=ye, noise=0.1), absorption=dict(spectral=v, brightness=ya, noise=0.01),
                n_sightlines=n_sl)
B = n_sl * per
th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal"]
theta = torch.from_numpy(eng.pack(synthetic.draw_free(kind, N, B, rng, thermal_fn=th))).cuda()
for order in ("grouped", "shuffled"):
    sl = np.repeat(np.arange(n_sl, dtype=np.int32), per)
    if order == "shuffled":
        rng.shuffle(sl)
    sl_d = torch.from_numpy(sl).cuda()
    lp = torch.empty(B, dtype=torch.float64, device="cuda"); gr = torch.empty_like(theta)
    for _ in range(7):
        eng.logp_grad_dev(B, theta, lp, gr, sightline=sl_d)
    eng.synchronize()
    st = eng.stage_ms(5).mean(axis=0)
    units = B * N * 2 * S
    print(f"{order}: B={B} map {st[0]:.3f} moments {st[1]:.3f} epilogue {st[3]:.3f} ms -> {units / st.sum() / 1e6:.1f} Gunits/s", flush=True)
