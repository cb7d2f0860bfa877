"""CPU, world_size = 2 over gloo: the N > 1 host logic (shard ranges, feature-table gather) -- the same code
path bench.py / a multi-GPU analysis uses with NCCL.  The per-shard compute is the oracle here (tests may use it)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nbodysimproject_b200.sharding import analyze_sharded, shard_range
    from oracle import nbody_oracle as O
    rng = np.random.RandomState(0)                     # every rank builds the same global inputs
    m = rng.uniform(0.5, 2.0, (n_total, 3))
    q = rng.randn(n_total, 3, 2)

    def compute(lo, hi):
        rows = np.empty((hi - lo, 2))
        for k, i in enumerate(range(lo, hi)):
            rows[k, 0] = O.softened_potential(q[i], m[i], 1.0, 0.05)
            rows[k, 1] = O.dV_d_epsilon(q[i], m[i], 0.05, 1.0)
        return rows

    full = analyze_sharded(compute, n_total)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full)
    lo, hi = shard_range(n_total, rank, world)
    np.save(os.path.join(out_dir, f"range{rank}.npy"), np.array([lo, hi]))
    dist.destroy_process_group()


def test_shard_range_partitions():
    from nbodysimproject_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_rank_gather_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp
    from oracle import nbody_oracle as O
    n_total, world = 11, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "rank0.npy")
    b = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b) and a.shape == (n_total, 2)
    r0, r1 = np.load(tmp_path / "range0.npy"), np.load(tmp_path / "range1.npy")
    assert r0[0] == 0 and r0[1] == r1[0] and r1[1] == n_total
    rng = np.random.RandomState(0)
    m = rng.uniform(0.5, 2.0, (n_total, 3))
    q = rng.randn(n_total, 3, 2)
    ref = np.array([[O.softened_potential(q[i], m[i], 1.0, 0.05), O.dV_d_epsilon(q[i], m[i], 0.05, 1.0)] for i in range(n_total)])
    assert np.array_equal(a, ref)        # identical whatever the rank count
