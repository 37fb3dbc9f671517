"""BASELINE configs[3] across GPUs (strong scaling): 10^4 sightlines x 5 clouds x 1000 channels,
`per` chains per sightline, sightlines sharded contiguously over the ranks
(caribou_hi_b200.sharding), per-step NCCL all_gather of the per-shard logp / gradient sums.
launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/survey_scale.py [per]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from caribou_hi_b200 import Engine, sharding, synthetic

per = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
kind, N, S, n_sl = "emission_absorption_physical", 5, 1000, 10_000
lo, hi = sharding.shard_bounds(n_sl, world, rank)
rng = np.random.default_rng(8 + rank)
v = np.linspace(-100.0, 100.0, S)
ye = rng.normal(10, 5, (hi - lo, S)); ya = rng.normal(0.3, 0.1, (hi - lo, S))
eng = Engine(kind, N, device=local, prior_amplitude=1e21)
eng.set_spectra(emission=dict(spectral=v, brightness=ye, noise=0.1), absorption=dict(spectral=v, brightness=ya, noise=0.01),
                n_sightlines=hi - lo)
B = (hi - lo) * per
th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal"]
dev = torch.device("cuda", local)
theta = torch.from_numpy(eng.pack(synthetic.draw_free(kind, N, B, rng, thermal_fn=th))).to(dev)
sl_d = torch.from_numpy(np.repeat(np.arange(hi - lo, dtype=np.int32), per)).to(dev)
lp = torch.empty(B, dtype=torch.float64, device=dev); gr = torch.empty_like(theta)
stream = torch.cuda.ExternalStream(eng.stream, device=dev)
gathered = torch.empty((world, 1 + 10 * N), dtype=torch.float64, device=dev)


def step():
    eng.logp_grad_dev(B, theta, lp, gr, sightline=sl_d)
    if world > 1:
        with torch.cuda.stream(stream):
            summary = torch.cat([torch.nan_to_num(lp).sum().reshape(1), torch.nan_to_num(gr).sum(dim=0).reshape(-1)])
            dist.all_gather_into_tensor(gathered, summary)


for _ in range(5):
    step()
eng.synchronize(); torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 20
e0.record(stream)
for _ in range(steps):
    step()
e1.record(stream)
eng.synchronize(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    units = n_sl * per * N * 2 * S
    print(f"survey n_gpus={world}: {n_sl} sightlines x {per} chains, {ms.item():.3f} ms per sweep -> {units / ms.item() / 1e6:.1f} Gunits/s", flush=True)
if world > 1:
    dist.destroy_process_group()
