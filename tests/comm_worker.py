"""Worker of tests/test_gpu_comm.py::test_nccl_two_ranks (one process per GPU under torchrun): every rank
evaluates its shard of a seeded batch, the library's communicator (chi_comm_*, NCCL) reduces / gathers the
shard sums, and every rank checks them against its own evaluation of the whole batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from caribou_hi_b200 import sharding  # noqa: E402
from helpers import engine_for, make_case  # noqa: E402


def main():
    rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # only ships the 128-byte id: the collectives are the library's
    spec, data, free, baseline = make_case("emission_absorption_physical", 3, 160, 160, 1001, seed=5, shared_axis=True)
    eng = engine_for(spec, data, 0, device=local)
    sharding.init_engine_comm(eng)
    assert eng.comm_size == ws and eng.comm_rank == rank
    rows = eng.pack_rows(free)
    ce, ca = baseline["baseline_emission_norm"], baseline["baseline_absorption_norm"]
    dev = torch.device("cuda", local)
    # reference: the whole batch on this GPU
    lp, gr, _, _ = eng.logp_grad_rows(rows, ce, ca)
    fin = np.isfinite(lp)
    se = sharding.ShardedEngine(eng)
    lo, hi = se.shard(rows.shape[0])
    assert (lo, hi) == sharding.shard_bounds(rows.shape[0], ws, rank)
    r_d = torch.from_numpy(rows[lo:hi]).to(dev)
    ce_d, ca_d = torch.from_numpy(ce[lo:hi]).to(dev), torch.from_numpy(ca[lo:hi]).to(dev)
    tot_lp, tot_g, n_bad = se.sums(r_d, ce_d, ca_d, reduce=True)
    assert n_bad == int((~fin).sum()), (n_bad, int((~fin).sum()))
    assert abs(tot_lp - lp[fin].sum()) <= 1e-12 * abs(lp[fin].sum())
    ref_g = gr[fin].sum(axis=0)
    assert np.all(np.abs(tot_g - ref_g) <= 1e-11 * np.abs(ref_g).max())
    table = se.sums(r_d, ce_d, ca_d, reduce=False)
    assert table.shape == (ws, 2 + ref_g.size)
    for r in range(ws):
        a, b = sharding.shard_bounds(rows.shape[0], ws, r)
        f = fin[a:b]
        assert abs(table[r, 0] - lp[a:b][f].sum()) <= 1e-12 * abs(lp[a:b][f].sum())
        assert table[r, -1] == (~f).sum()
    assert abs(table[:, 0].sum() - tot_lp) <= 1e-12 * abs(tot_lp)
    eng.comm_destroy()
    eng.close()
    dist.barrier()
    if rank == 0:
        print("COMM_WORKER_OK", ws)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
