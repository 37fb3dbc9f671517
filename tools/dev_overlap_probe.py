"""CHI_OPT_DEV_OVERLAP: consecutive device-resident calls on one lane (default) against two alternating lanes
(map / epilogue of one call next to the channel kernel of the other), configs[1], device-timed; same bits."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from caribou_hi_b200 import Engine

eng = Engine(bench.KIND, bench.N_CLOUDS, prior_amplitude=1e21)
thetas = bench.build_inputs(eng, bench.DRAWS, 0, 4)
dev = torch.device("cuda", 0)
B, N = bench.DRAWS, bench.N_CLOUDS
th = [torch.from_numpy(t).to(dev) for t in thetas]
lp = [torch.empty(B, dtype=torch.float64, device=dev) for _ in th]
gr = [torch.empty((B, 10, N), dtype=torch.float64, device=dev) for _ in th]
ce = torch.zeros((B, 1), dtype=torch.float64, device=dev); ca = torch.zeros_like(ce); ge = torch.empty_like(ce); ga = torch.empty_like(ce)
stream = torch.cuda.ExternalStream(eng.stream)
ref = None
for mode in (0, 1, 0, 1):
    eng.set_option("dev_overlap", mode)
    def step(i):
        k = i % len(th)
        eng.logp_grad_dev(B, th[k], lp[k], gr[k], coeff_e=ce, coeff_a=ca, gcoeff_e=ge, gcoeff_a=ga)
    for i in range(4): step(i)
    eng.dev_join(); eng.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(60): step(i)
    eng.dev_join()
    e1.record(stream); eng.synchronize()
    ms = e0.elapsed_time(e1) / 60
    out = (lp[1].cpu().numpy().copy(), gr[1].cpu().numpy().copy())
    same = "" if ref is None else f" same bits as mode 0: {np.array_equal(out[0], ref[0], equal_nan=True) and np.array_equal(out[1], ref[1], equal_nan=True)}"
    if ref is None: ref = out
    print(f"dev_overlap={mode}: {ms:.4f} ms per step -> {B*N*2000/ms/1e6:.1f} Gunits/s{same}", flush=True)
