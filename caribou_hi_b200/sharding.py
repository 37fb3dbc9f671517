"""Multi-GPU layer: one process per GPU (``torchrun``), the batch of independent
evaluations (draws x chains x sightlines) split contiguously across ranks.

The arithmetic has no exchange step: every draw is independent, so shards never talk to
each other while computing; results are only *collected*.  On GPUs the library owns the NCCL
communicator and the reductions (``chi_comm_*``, ``chi_shard_sums_dev``, csrc/comm.cuh):
:func:`init_engine_comm` creates it from the ``torch.distributed`` world (used only to ship the
128-byte id) and :class:`ShardedEngine` is the device-resident front end.  The functions below it
are the same logic on ``torch.distributed`` tensors (gloo on CPU in the tests):

* ``gather_rows``      all-gather per-draw ``logp[B_local]`` / ``grad[B_local, ...]`` back
                       into batch order (uneven shards are padded for the collective)
* ``reduce_sum``       all-reduce of per-shard sums (survey-total logp / gradient)

Posterior-predictive spectra stay sharded on purpose: 1e6 draws x 8192 channels x 8 B is
65 GB, far more than any consumer wants on one rank.
"""

import numpy as np


def shard_bounds(n_items, world_size, rank):
    """Contiguous split: ranks ``< n_items % world_size`` get one extra item."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size/rank")
    base, extra = divmod(int(n_items), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_items, world_size):
    return [shard_bounds(n_items, world_size, r)[1] - shard_bounds(n_items, world_size, r)[0]
            for r in range(world_size)]


def shard_by_sightline(sightline, world_size, rank):
    """Survey sharding (BASELINE configs[3]): all draws of a sightline go to the rank that
    owns the sightline, so each GPU only needs its own sightlines' spectra.  Returns the
    indices of the draws owned by ``rank`` (ascending) and the owned sightline range."""
    sightline = np.asarray(sightline)
    n_sl = int(sightline.max()) + 1 if sightline.size else 0
    lo, hi = shard_bounds(n_sl, world_size, rank)
    idx = np.nonzero((sightline >= lo) & (sightline < hi))[0]
    return idx, (lo, hi)


def _dist():
    import torch.distributed as dist

    return dist


def world():
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def gather_rows(local, n_total, device=None):
    """All-gather row blocks of a contiguous split back into one ``[n_total, ...]`` tensor on
    every rank.  ``local`` is a torch tensor (CUDA with NCCL, CPU with gloo) holding this
    rank's rows."""
    import torch

    dist = _dist()
    ws, rank = world()
    if ws == 1:
        return local
    sizes = shard_sizes(n_total, ws)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank}: expected {sizes[rank]} rows, got {local.shape[0]}")
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: sizes[rank]] = local
    out = torch.empty((ws * pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf)
    parts = [out[r * pad : r * pad + sizes[r]] for r in range(ws)]
    return torch.cat(parts, dim=0)


def reduce_sum(local):
    """All-reduce(sum) of a per-shard partial (e.g. survey-total logp and gradient)."""
    dist = _dist()
    ws, _ = world()
    if ws > 1:
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
    return local


class ShardedLikelihood:
    """Evaluates this rank's shard of a batch on its own GPU and collects the results.

    ``evaluate`` is any callable ``theta_local -> (logp_local, grad_local)`` on NumPy arrays
    (normally ``lambda th: engine.logp_grad(th)[:2]``); keeping it a callable lets the CPU
    test-suite drive the sharding/collection logic under gloo without a GPU.
    """

    def __init__(self, evaluate, device=None):
        self.evaluate = evaluate
        self.device = device
        self.n_bad = 0  # non-finite draws (over all ranks) seen by the last total_logp_grad

    def logp_grad(self, theta):
        """``theta [B, ...]`` is the FULL batch, identical on every rank; returns the full
        ``(logp [B], grad [B, ...])`` on every rank."""
        import torch

        ws, rank = world()
        B = theta.shape[0]
        lo, hi = shard_bounds(B, ws, rank)
        logp, grad = self.evaluate(theta[lo:hi])
        if ws == 1:
            return logp, grad
        dev = self.device if self.device is not None else "cpu"
        lp = gather_rows(torch.as_tensor(logp).to(dev), B)
        gr = gather_rows(torch.as_tensor(grad).to(dev), B)
        return lp.cpu().numpy(), gr.cpu().numpy()

    def total_logp_grad(self, theta):
        """Sum over the whole batch (hierarchical / survey-total use): only one scalar and
        one gradient block cross the interconnect."""
        import torch

        ws, rank = world()
        lo, hi = shard_bounds(theta.shape[0], ws, rank)
        logp, grad = self.evaluate(theta[lo:hi])
        dev = self.device if self.device is not None else "cpu"
        # plain sums: a NaN draw makes the total NaN, as it does in the reference (where the sampler
        # then rejects the point); the third value counts this shard's non-finite draws for callers
        # that want to know why
        bad = np.count_nonzero(~np.isfinite(logp))
        packed = torch.cat([torch.as_tensor(np.sum(logp, keepdims=True)).reshape(1),
                            torch.as_tensor(np.sum(grad, axis=0)).reshape(-1),
                            torch.as_tensor([float(bad)], dtype=torch.float64)]).to(dev)
        packed = reduce_sum(packed).cpu().numpy()
        self.n_bad = int(packed[-1])
        return packed[0], packed[1:-1].reshape(grad.shape[1:])


# ---- the library's own communicator (NCCL inside the handle) --------------------------------------
def init_engine_comm(engine):
    """Create ``engine``'s NCCL communicator for the current ``torch.distributed`` world: rank 0 makes
    the id (``chi_comm_unique_id``), the process group ships it, every rank calls ``chi_comm_init``.
    A world of one needs no communicator (the collectives degenerate to copies)."""
    ws, rank = world()
    if ws == 1:
        return
    dist = _dist()
    box = [type(engine).comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    engine.comm_init(box[0], rank, ws)


class ShardedEngine:
    """This rank's shard of a batch on its own GPU, results collected by the library: per-shard sums of
    logp and gradient (``chi_shard_sums_dev``) gathered or reduced over the ranks with NCCL on the
    handle's communication stream (``chi_allgather_dev`` / ``chi_allreduce_sum_dev``).  Tensors are
    ``torch`` CUDA tensors on the engine's device (device memory is all torch is used for)."""

    def __init__(self, engine):
        import torch

        self.engine = engine
        self.torch = torch
        self.device = torch.device("cuda", engine.device)
        self._buf = {}

    def _tensor(self, key, shape):
        t = self._buf.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = self.torch.empty(shape, dtype=self.torch.float64, device=self.device)
            self._buf[key] = t
        return t

    def shard(self, n_items):
        """``(start, stop)`` of this rank's contiguous shard of ``n_items``."""
        return shard_bounds(n_items, self.engine.comm_size, self.engine.comm_rank)

    def sums(self, rows_dev, coeff_e=None, coeff_a=None, sightline=None, reduce=True):
        """Evaluate the local packed rows ``[B_local, n_rows, N]`` (device) and return, on the host,
        ``(sum logp, sum grad [n_rows, N], non-finite draws)`` over ALL ranks (``reduce=True``:
        all-reduce) or the per-rank table ``[n_ranks, 2 + n_rows N]`` (``reduce=False``: all-gather)."""
        eng = self.engine
        B, n_rows, N = rows_dev.shape
        width = n_rows * N
        lp = self._tensor("lp", (B,))
        gr = self._tensor("gr", (B, n_rows, N))
        gce = self._tensor("gce", (B, max(eng.nb_e, 1)))
        gca = self._tensor("gca", (B, max(eng.nb_a, 1)))
        eng.logp_grad_rows_dev(B, rows_dev, lp, gr, coeff_e=coeff_e, coeff_a=coeff_a, sightline=sightline,
                               gcoeff_e=gce if eng.nb_e else None, gcoeff_a=gca if eng.nb_a else None)
        s = self._tensor("sums", (width + 2,))
        eng.shard_sums_dev(B, width, lp, gr, s)
        if reduce:
            eng.allreduce_sum_dev(s, width + 2)
            eng.comm_synchronize()
            h = s.cpu().numpy()
            return float(h[0]), h[1 : 1 + width].reshape(n_rows, N), int(round(h[1 + width]))
        table = self._tensor("table", (eng.comm_size, width + 2))
        eng.allgather_dev(s, table, width + 2)
        eng.comm_synchronize()
        return table.cpu().numpy()
