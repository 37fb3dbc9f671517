"""Quick device-side timing of the stage kernels for a few shapes (development aid).
usage: python tools/quick_bench.py [draws]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caribou_hi_b200 import Engine, synthetic

def run(kind, N, S_e, S_a, B, lorentzian, reps=5):
    eng = Engine(kind, N, lorentzian=lorentzian, prior_fwhm_L=1.0 if lorentzian else 0.0, prior_amplitude=1e21)
    ve = np.linspace(-100, 100, S_e); va = np.linspace(-100, 100, S_a)
    rng = np.random.default_rng(0)
    eng.set_spectra(emission=dict(spectral=ve, brightness=rng.normal(10, 5, S_e), noise=0.1, basis=np.full((1, S_e), 0.1)),
                    absorption=dict(spectral=va, brightness=rng.normal(0.3, 0.1, S_a), noise=0.01, basis=np.full((1, S_a), 0.01)))
    th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal"]
    free = synthetic.draw_free(kind, N, B, rng, lorentzian=lorentzian, thermal_fn=th)
    theta = torch.from_numpy(eng.pack(free)).cuda()
    lp = torch.empty(B, dtype=torch.float64, device="cuda"); gr = torch.empty((B, 10, N), dtype=torch.float64, device="cuda")
    ce = torch.zeros((B, 1), dtype=torch.float64, device="cuda"); ca = torch.zeros((B, 1), dtype=torch.float64, device="cuda")
    ge = torch.empty_like(ce); ga = torch.empty_like(ca)
    torch.cuda.synchronize()
    for _ in range(reps + 2):
        eng.logp_grad_dev(B, theta, lp, gr, coeff_e=ce, coeff_a=ca, gcoeff_e=ge, gcoeff_a=ga)
    eng.synchronize()
    st = eng.stage_ms(reps).mean(axis=0)
    units = B * N * (S_e + S_a)
    tot = st.sum()
    print(f"{kind} N={N} S={S_e}+{S_a} B={B} pv={int(lorentzian)}: map {st[0]:.3f} em {st[1]:.3f} abs {st[2]:.3f} epi {st[3]:.3f} ms"
          f" | total {tot:.3f} ms -> {units / tot / 1e6:.1f} Gunits/s (em {B*N*S_e/st[1]/1e6:.1f}, abs {B*N*S_a/st[2]/1e6:.1f})", flush=True)
    eng.close()

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    kind = "emission_absorption_physical"
    run(kind, 4, 1000, 1000, B, False)
    run(kind, 4, 1000, 1000, B, True)
    if os.environ.get("QUICK"): sys.exit(0)
    run(kind, 8, 2048, 2048, max(B // 8, 64), True)
    run(kind, 8, 2048, 2048, 64, True, reps=20)
    run(kind, 2, 334, 166, B, False)
    run(kind, 5, 1000, 1000, B // 2, False)
    run(kind, 1, 1000, 1000, B, False)
