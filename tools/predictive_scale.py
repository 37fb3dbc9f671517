"""BASELINE configs[4] across GPUs: posterior-predictive generation, 10^6 draws x 8 clouds x
4096 + 4096 channels of emission and absorption spectra (pseudo-Voigt), draws sharded contiguously
over the ranks (caribou_hi_b200.sharding).  The spectra stay sharded (65 GB in fp64 at full size);
what crosses NVLink per pass is the posterior-predictive MEAN spectrum of each dataset (an NCCL
all-reduce of 2 x 4096 sums).  Times the fp64 kernels and the optional fp32 path.
launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/predictive_scale.py [total_draws]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from caribou_hi_b200 import Engine, sharding, synthetic

total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
kind, N, S = "emission_absorption_physical", 8, 4096
lo, hi = sharding.shard_bounds(total, world, rank)
B = hi - lo
dev = torch.device("cuda", local)
eng = Engine(kind, N, device=local, lorentzian=True, prior_fwhm_L=1.0, prior_amplitude=1e21)
v = np.linspace(-100.0, 100.0, S)
eng.set_spectra(emission=dict(spectral=v, brightness=np.zeros(S), noise=0.1, basis=np.full((1, S), 0.1)),
                absorption=dict(spectral=v, brightness=np.zeros(S), noise=0.01, basis=np.full((1, S), 0.01)))
th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal"]
rng = np.random.default_rng(100 + rank)
# the draws of a shard: 16384 distinct prior draws tiled to the shard size (host RNG time, not the point here)
base = eng.pack(synthetic.draw_free(kind, N, min(B, 16384), rng, lorentzian=True, thermal_fn=th))
theta = torch.from_numpy(base).to(dev).repeat((B + base.shape[0] - 1) // base.shape[0], 1, 1)[:B].contiguous()
stream = torch.cuda.ExternalStream(eng.stream, device=dev)


def run(dtype, label):
    mu_e = torch.empty((B, S), dtype=dtype, device=dev); mu_a = torch.empty((B, S), dtype=dtype, device=dev)
    mean = torch.empty((2, S), dtype=torch.float64, device=dev)
    call = eng.spectra_dev if dtype == torch.float64 else eng.spectra_f32_dev

    def step():
        call(B, theta, mu_e, mu_a)
        with torch.cuda.stream(stream):
            mean[0] = torch.nan_to_num(mu_e).sum(dim=0, dtype=torch.float64)
            mean[1] = torch.nan_to_num(mu_a).sum(dim=0, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(mean)

    for _ in range(2):
        step()
    eng.synchronize(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    steps = 5
    e[0].record(stream)
    for _ in range(steps):
        call(B, theta, mu_e, mu_a)
    e[1].record(stream)
    for _ in range(steps):
        step()
    e[2].record(stream)
    eng.synchronize(); torch.cuda.synchronize()
    ms = torch.tensor([e[0].elapsed_time(e[1]) / steps, e[1].elapsed_time(e[2]) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        units = total * N * 2 * S
        k, full = ms.tolist()
        print(f"predictive {label} n_gpus={world}: {total} draws ({B} per GPU, {2 * B * S * mu_e.element_size() / 1e9:.1f} GB of spectra per GPU): "
              f"kernels {k:.2f} ms -> {units / k / 1e6:.0f} Gunits/s; with the posterior-predictive mean (column sums + NCCL all-reduce) "
              f"{full:.2f} ms -> {units / full / 1e6:.0f} Gunits/s; mean spectrum checksum {float(mean.sum()) / total:.6e}", flush=True)
    del mu_e, mu_a


run(torch.float64, "fp64")
run(torch.float32, "fp32")
if world > 1:
    dist.destroy_process_group()
