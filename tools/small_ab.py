"""Small-batch evaluation: one-launch kernel (csrc/small.cuh) against the three-launch pipeline, a few
shapes and tuning settings, device-timed back to back (CUDA events on the handle's stream).
usage: python tools/small_ab.py [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caribou_hi_b200 import Engine, synthetic

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50


def case(kind, N, S, B, lorentzian, shared=True, model=False):
    eng = Engine(kind, N, lorentzian=lorentzian, prior_fwhm_L=1.0 if lorentzian else 0.0, prior_amplitude=1e21 if "physical" in kind else 50.0)
    v = np.linspace(-100, 100, S)
    va = v if shared else np.linspace(-50, 50, S // 2)
    rng = np.random.default_rng(0)
    eng.set_spectra(emission=dict(spectral=v, brightness=rng.normal(10, 5, S), noise=0.1, basis=np.full((1, S), 0.1)),
                    absorption=dict(spectral=va, brightness=rng.normal(0.3, 0.1, va.size), noise=0.01, basis=np.full((1, va.size), 0.01)))
    th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal" if "physical" in kind else "tkin"]
    theta = torch.from_numpy(eng.pack(synthetic.draw_free(kind, N, B, rng, lorentzian=lorentzian, thermal_fn=th))).cuda()
    lp = torch.empty(B, dtype=torch.float64, device="cuda"); gr = torch.empty((B, 10, N), dtype=torch.float64, device="cuda")
    ce = torch.zeros((B, 1), dtype=torch.float64, device="cuda"); ca = torch.zeros((B, 1), dtype=torch.float64, device="cuda")
    ge = torch.empty_like(ce); ga = torch.empty_like(ca)
    stream = torch.cuda.ExternalStream(eng.stream)
    out = {}
    for label, path, nl, chunks in (("pipeline", 1, 0, 0), ("one-launch auto", 2, 0, 0), ("one-launch NL=1", 2, 1, 0),
                                    ("one-launch NL=2", 2, 2, 0), ("one-launch NL=3", 2, 3, 0), ("one-launch NL=4", 2, 4, 0)):
        if nl > 1 and N < nl:
            continue
        eng.set_option("small_path", path); eng.set_option("small_nl", nl)
        auto = max(1, (148 * 2) // B)
        eng.set_option("small_chunks", 0 if chunks == 0 else (max(1, auto // 2) if chunks == -2 else auto * 2))
        for _ in range(3):
            eng.logp_grad_dev(B, theta, lp, gr, coeff_e=ce, coeff_a=ca, gcoeff_e=ge, gcoeff_a=ga)
        eng.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            eng.logp_grad_dev(B, theta, lp, gr, coeff_e=ce, coeff_a=ca, gcoeff_e=ge, gcoeff_a=ga)
        e1.record(stream); eng.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        # the same evaluation replayed from a CUDA graph (what the sampler does): no host launch cost
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for _ in range(20):
                eng.logp_grad_dev(B, theta, lp, gr, coeff_e=ce, coeff_a=ca, gcoeff_e=ge, gcoeff_a=ga)
        with torch.cuda.stream(stream):
            g.replay(); eng.synchronize()
            e0.record(stream)
            for _ in range(5):
                g.replay()
            e1.record(stream)
        eng.synchronize()
        us_g = e0.elapsed_time(e1) / 100 * 1e3
        res = (lp.cpu().numpy().copy(), gr.cpu().numpy().copy())
        if "ref" not in out:
            out["ref"] = res
        fin = np.isfinite(out["ref"][0])
        dl = np.max(np.abs(res[0][fin] - out["ref"][0][fin]) / (np.abs(out["ref"][0][fin]) + 1e-300)) if fin.any() else 0
        scale = np.max(np.abs(out["ref"][1][fin]), axis=(1, 2), keepdims=True) if fin.any() else 1
        dg = np.max(np.abs(res[1][fin] - out["ref"][1][fin]) / scale) if fin.any() else 0
        same_nan = np.array_equal(np.isfinite(res[0]), fin)
        if path == 2 and nl == 0 and chunks == 0:
            st = torch.zeros(16, dtype=torch.int64, device="cuda")
            eng.set_option("small_stamps", st.data_ptr())
            eng.logp_grad_dev(B, theta, lp, gr, coeff_e=ce, coeff_a=ca, gcoeff_e=ge, gcoeff_a=ga); eng.synchronize()
            eng.set_option("small_stamps", 0)
            c = st.cpu().numpy() / 1.965e3
            print(f"    block 0 phases (us at 1.965 GHz): map {c[0]:.1f} chan-warp0 done {c[2]:.1f} | jac-warp: jac {c[9]:.1f} | block joined {c[3]:.1f} record+counter {c[4]:.1f} logp {c[5]:.1f} grad {c[6]:.1f}")
        print(f"  {label:22s} {us:7.1f} us/eval host loop, {us_g:6.1f} in a graph  {B*N*(S+va.size)/us_g/1e3:7.1f} Gunits/s  max rel dlogp {dl:.1e} dgrad {dg:.1e} nan-pattern {'same' if same_nan else 'DIFFERS'}", flush=True)
    eng.close()


for kind, N, S, B, lor, shared in (("emission_absorption_physical", 8, 2048, 64, True, True),
                                   ("emission_absorption", 2, 334, 8, False, True),
                                   ("emission_absorption", 2, 334, 8, False, False),
                                   ("emission_absorption_physical", 4, 1000, 64, False, True),
                                   ("emission_absorption_physical", 4, 1000, 512, False, True),
                                   ("emission_absorption_physical", 4, 1000, 2048, False, True),
                                   ("emission_absorption_physical", 8, 2048, 512, True, True),
                                   ("emission_absorption_physical", 5, 1000, 256, True, False),
                                   ("emission_absorption_physical", 16, 512, 32, True, True)):
    print(f"{kind} N={N} S={S} B={B} pv={int(lor)} shared={int(shared)}", flush=True)
    case(kind, N, S, B, lor, shared)
