"""CPU, world_size 2, gloo: the N>1 host path -- contiguous sharding by global replication index and the single
allreduce of the {sum x, sum x^2, n} vector -- gives the same statistics as one rank."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from manusim_b200.experiment import generate_run_seeds, shard_bounds, stats_from_moments


def _moments(x):
    """host restatement of msim_reduce_metrics for pre-computed per-replication metric values [n, K]"""
    ok = ~np.isnan(x)
    xz = np.where(ok, x, 0.0)
    return np.stack([xz.sum(0), (xz * xz).sum(0), ok.sum(0).astype(np.float64)], axis=1)


def _worker(rank, world, port, n, K, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    x = rng.gamma(2.0, 1.5, size=(n, K))
    x[rng.random((n, K)) < 0.05] = np.nan
    seeds = generate_run_seeds(123, n)
    lo, hi = shard_bounds(n, world, rank)
    assert seeds[lo:hi] == generate_run_seeds(123, n)[lo:hi]
    mom = torch.from_numpy(_moments(x[lo:hi]))
    dist.all_reduce(mom)
    if rank == 0:
        np.save(out, mom.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_moments_allreduce(tmp_path):
    n, K, world = 1001, 37, 2
    out = str(tmp_path / "mom.npy")
    mp.spawn(_worker, args=(world, _free_port(), n, K, out), nprocs=world, join=True)
    got = np.load(out)
    rng = np.random.default_rng(11)
    x = rng.gamma(2.0, 1.5, size=(n, K))
    x[rng.random((n, K)) < 0.05] = np.nan
    want = _moments(x)
    np.testing.assert_allclose(got, want, rtol=1e-12)
    st = stats_from_moments(got)
    np.testing.assert_allclose(st["mean"], np.nanmean(x, axis=0), rtol=1e-10)
    np.testing.assert_allclose(st["std"], np.nanstd(x, axis=0, ddof=1), rtol=1e-8)
