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
    # single precision (chi_spectra_f32_dev): same draws, float32 outputs
    e32 = torch.empty((B, S), dtype=torch.float32, device="cuda"); a32 = torch.empty((B, S), dtype=torch.float32, device="cuda")
    for _ in range(2): eng.spectra_f32_dev(B, theta, e32, a32)
    eng.synchronize()
    e0.record(stream)
    for _ in range(5): eng.spectra_f32_dev(B, theta, e32, a32)
    e1.record(stream); eng.synchronize()
    ms = e0.elapsed_time(e1) / 5
    def worst(x32, x64):
        ok = torch.isfinite(x64).all(dim=1)
        scale = x64[ok].abs().amax(dim=1, keepdim=True)
        return float(((x32[ok].double() - x64[ok]).abs() / scale).max())
    print(f"  f32: {ms:.3f} ms -> {units/ms/1e6:.1f} Gunits/s, output {2*B*S*4/ms/1e6:.0f} GB/s; worst |f32-f64|/max|f64| "
          f"emission {worst(e32, mu_e):.2e} absorption {worst(a32, mu_a):.2e}", flush=True)
    # predictive draws (spectra + counter-based Normal noise), fp64 and fp32
    for f32, (oe, oa) in ((False, (mu_e, mu_a)), (True, (e32, a32))):
        for _ in range(2): eng.predictive_dev(B, theta, oe, oa, seed=1, f32=f32)
        eng.synchronize()
        e0.record(stream)
        for _ in range(5): eng.predictive_dev(B, theta, oe, oa, seed=1, f32=f32)
        e1.record(stream); eng.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"  predictive draws {'f32' if f32 else 'f64'}: {ms:.3f} ms -> {units/ms/1e6:.1f} Gunits/s", flush=True)
    eng.close()

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "c5":  # profiling target: BASELINE configs[4] per-GPU shape only
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 25000, 8, 4096, True)
        sys.exit(0)
    run(25000, 8, 4096, True)
    run(25000, 8, 4096, False)
    run(100000, 4, 1000, False)
