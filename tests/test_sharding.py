"""Host-side multi-GPU logic under gloo, world_size 2, on CPU (no GPU needed): the shard
split, the padded all-gather back into batch order and the survey-total all-reduce."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from caribou_hi_b200 import sharding


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 100_000, 10**6 + 3):
        for ws in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            assert sizes == sharding.shard_sizes(n, ws)
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def test_shard_by_sightline():
    rng = np.random.default_rng(0)
    sl = rng.integers(0, 10, size=200)
    seen = []
    for r in range(4):
        idx, (lo, hi) = sharding.shard_by_sightline(sl, 4, r)
        assert np.all((sl[idx] >= lo) & (sl[idx] < hi))
        seen.append(idx)
    assert sorted(np.concatenate(seen).tolist()) == list(range(200))


def _fake_evaluate(theta):
    """Stands in for engine.logp_grad on a shard: any per-row function of theta."""
    logp = -0.5 * (theta**2).sum(axis=(1, 2))
    grad = -theta
    return logp, grad


def _worker(rank, world_size, port, B, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        rng = np.random.default_rng(42)  # same full batch on every rank
        theta = rng.standard_normal((B, 10, 3))
        sh = sharding.ShardedLikelihood(_fake_evaluate)
        logp, grad = sh.logp_grad(theta)
        ref_lp, ref_g = _fake_evaluate(theta)
        assert logp.shape == (B,) and grad.shape == theta.shape
        np.testing.assert_array_equal(logp, ref_lp)
        np.testing.assert_array_equal(grad, ref_g)
        tot, gtot = sh.total_logp_grad(theta)
        np.testing.assert_allclose(tot, ref_lp.sum(), rtol=1e-13)
        np.testing.assert_allclose(gtot, ref_g.sum(axis=0), rtol=1e-12, atol=1e-12)
        # raw gather of uneven shards
        lo, hi = sharding.shard_bounds(B, world_size, rank)
        rows = sharding.gather_rows(torch.arange(lo, hi, dtype=torch.float64), B)
        assert torch.equal(rows, torch.arange(B, dtype=torch.float64))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("B", [7, 64])
def test_gloo_world_size_2(tmp_path, B):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
