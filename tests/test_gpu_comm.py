"""The library's multi-GPU layer (include/chi.h "multi-GPU", csrc/comm.cuh): shard sums on the device and
the NCCL collectives behind the handle.  The two-rank test needs two GPUs (`gpurun --gpus 2`)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import engine_for, make_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_sums_match_numpy_and_skip_nan_draws(cuda_device):
    import torch

    spec, data, free, baseline = make_case("emission_absorption_physical", 4, 200, 200, 5000, seed=21, shared_axis=True)
    eng = engine_for(spec, data, 0)
    rows = eng.pack_rows(free)
    ce, ca = baseline["baseline_emission_norm"], baseline["baseline_absorption_norm"]
    lp, gr, _, _ = eng.logp_grad_rows(rows, ce, ca)
    fin = np.isfinite(lp)
    assert 0 < (~fin).sum() < lp.size  # the prior of the physical models produces NaN draws, as the reference does
    dev = torch.device("cuda", 0)
    lp_d, gr_d = torch.from_numpy(lp).to(dev), torch.from_numpy(gr).to(dev)
    width = gr.shape[1] * gr.shape[2]
    out = [torch.zeros(width + 2, dtype=torch.float64, device=dev) for _ in range(2)]
    for o in out:
        eng.shard_sums_dev(lp.size, width, lp_d, gr_d, o)
    eng.comm_synchronize()
    a, b = out[0].cpu().numpy(), out[1].cpu().numpy()
    assert np.array_equal(a, b)  # fixed summation order: deterministic
    assert a[-1] == (~fin).sum()
    assert abs(a[0] - lp[fin].sum()) <= 1e-13 * abs(lp[fin].sum())
    ref = gr[fin].sum(axis=0).reshape(-1)
    assert np.all(np.abs(a[1:-1] - ref) <= 1e-12 * np.abs(ref).max())
    # one rank, no communicator: the gather is a copy, the reduction the identity
    rec = torch.zeros_like(out[0])
    eng.allgather_dev(out[0], rec, width + 2)
    eng.allreduce_sum_dev(out[1], width + 2)
    eng.comm_join()
    eng.synchronize()
    assert np.array_equal(rec.cpu().numpy(), a) and np.array_equal(out[1].cpu().numpy(), a)
    assert eng.comm_size == 1 and eng.comm_rank == 0
    # edge: a shard with no draws
    eng.shard_sums_dev(0, width, lp_d, gr_d, rec)
    eng.comm_synchronize()
    assert not rec.cpu().numpy().any()
    eng.close()


def test_nccl_two_ranks(cuda_device):
    """Two processes, two GPUs: chi_comm_init from an id shipped over a gloo group, shard sums reduced and
    gathered with the library's NCCL calls (tests/comm_worker.py)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "comm_worker.py")]
    env = dict(os.environ, NCCL_DEBUG="WARN")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0 and "COMM_WORKER_OK 2" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


@pytest.mark.parametrize("B", [37, 20000])
def test_host_sums_entry_points(cuda_device, B):
    """chi_logp_sums_rows[_async]: logp per draw + totals of the gradient over the finite draws; the
    per-draw gradient stays on the device.  Small batch (one piece) and a pipelined one (pieces on three
    lanes, sums added in piece order)."""
    from caribou_hi_b200 import PinnedArray

    spec, data, free, baseline = make_case("emission_absorption_physical", 3, 120, 120, B, seed=31, shared_axis=True)
    eng = engine_for(spec, data, 0)
    rows = eng.pack_rows(free)
    ce, ca = baseline["baseline_emission_norm"], baseline["baseline_absorption_norm"]
    lp, gr, _, _ = eng.logp_grad_rows(rows, ce, ca)
    fin = np.isfinite(lp)
    lp2, sums = eng.logp_sums_rows(rows, ce, ca)
    assert np.array_equal(lp2, lp, equal_nan=True)
    tot, g, n_bad = eng.unpack_sums(sums)
    assert n_bad == (~fin).sum()
    assert abs(tot - lp[fin].sum()) <= 1e-12 * abs(lp[fin].sum())
    ref = eng.unpack_rows(gr[fin].sum(axis=0, keepdims=True))
    for k in ref:
        assert np.all(np.abs(g[k] - ref[k][0]) <= 1e-11 * np.abs(gr[fin].sum(axis=0)).max()), k
    # the sums alone; twice the same bits (fixed order)
    none, sums_b = eng.logp_sums_rows(rows, ce, ca, want_logp=False)
    assert none is None and np.array_equal(sums_b, sums)
    # asynchronous, two in flight
    pin = {k: PinnedArray(v.shape) for k, v in (("rows", rows), ("ce", ce), ("ca", ca))}
    pin["rows"].array[...] = rows; pin["ce"].array[...] = ce; pin["ca"].array[...] = ca
    outs = [{"logp": PinnedArray((B,)).array, "sums": np.zeros_like(sums)}, {"sums": np.zeros_like(sums)}]
    t0 = eng.logp_sums_rows_async(pin["rows"].array, pin["ce"].array, pin["ca"].array, outs[0])
    t1 = eng.logp_sums_rows_async(pin["rows"].array, pin["ce"].array, pin["ca"].array, outs[1])
    eng.wait(t1)
    # the asynchronous call cuts the batch into other pieces than the blocking one: the same
    # bits between asynchronous calls, rounding-level agreement with the blocking call
    assert np.array_equal(outs[0]["sums"], outs[1]["sums"])
    assert np.all(np.abs(outs[0]["sums"] - sums) <= 1e-12 * np.abs(sums).max())
    assert np.array_equal(outs[0]["logp"], lp, equal_nan=True)
    eng.wait(t0)
    eng.close()


def test_overlapped_device_calls_give_the_same_bits(cuda_device):
    """CHI_OPT_DEV_OVERLAP: consecutive device-resident calls alternate between two internal lanes (map and
    epilogue on the lane's priority stream) and are ordered into the handle's stream by chi_dev_join."""
    import torch

    spec, data, free, baseline = make_case("emission_absorption_physical", 3, 160, 160, 8000, seed=41, shared_axis=True)
    eng = engine_for(spec, data, 0)
    dev = torch.device("cuda", 0)
    thetas = [torch.from_numpy(eng.pack({k: np.roll(v, s, axis=0) for k, v in free.items()})).to(dev) for s in (0, 1, 2)]
    ce = torch.from_numpy(baseline["baseline_emission_norm"]).to(dev)
    ca = torch.from_numpy(baseline["baseline_absorption_norm"]).to(dev)
    B = 8000
    res = {}
    for mode in (0, 1):
        eng.set_option("dev_overlap", mode)
        lp = [torch.zeros(B, dtype=torch.float64, device=dev) for _ in thetas]
        gr = [torch.zeros((B, 10, 3), dtype=torch.float64, device=dev) for _ in thetas]
        for rep in range(2):
            for k, th in enumerate(thetas):
                eng.logp_grad_dev(B, th, lp[k], gr[k], coeff_e=ce, coeff_a=ca)
        eng.dev_join()
        eng.synchronize()
        res[mode] = [(a.cpu().numpy(), b.cpu().numpy()) for a, b in zip(lp, gr)]
    for (a0, b0), (a1, b1) in zip(res[0], res[1]):
        assert np.isfinite(a0).any()
        assert np.array_equal(a0, a1, equal_nan=True) and np.array_equal(b0, b1, equal_nan=True)
    eng.close()
