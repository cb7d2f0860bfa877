"""Multi-GPU consistency check, run under torchrun (one rank per GPU, NCCL):
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_mgpu.py
(1) ensembles sharded by system + feature gather == the single-GPU feature table, bit for bit;
(2) large-N force with i-blocks sharded + in-place NCCL position all-gather == single-GPU force;
(3) large-N ham_soft Strang sub-steps sharded == single GPU (fp32 noise only: the fp64 atomics commute differently).
Exit code 0 = all consistent."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodysimproject_b200 import ensemble as E
from nbodysimproject_b200 import largen as LN
from nbodysimproject_b200 import sharding as S
from nbodysimproject_b200.generators import EnsembleInputs


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    # (1) ensemble: same global inputs on every rank, each analyses its shard
    rng = np.random.default_rng(123)
    buckets = EnsembleInputs.diverse(rng, 4096, n_max=8)
    for N, (m, q, v, soft, cohort) in sorted(buckets.items()):
        B = m.shape[0]
        rr, rv = rng.standard_normal((B, N, 2)), rng.standard_normal((B, N, 2))

        def compute(lo, hi):
            r = E.analyze_bucket(m[lo:hi], q[lo:hi], v[lo:hi], soft[lo:hi], 1.0, "yoshida4", 200, 0.01, "full",
                                 rr[lo:hi], rv[lo:hi], device=dev)
            return np.concatenate([r.dyn, r.static], axis=1)

        full = S.analyze_sharded(compute, B)
        if rank == 0:
            ref = compute(0, B)
            same = np.array_equal(full, ref, equal_nan=True)
            print(f"[ensemble N={N} B={B}] sharded == single: {same}", flush=True)
            ok &= bool(same)
    # (2) large-N force
    n = 1 << 15
    m, q, v = LN.make_disc(n, seed=4)
    sh = LN.LargeNSimulation(m, q, v, softening=1e-3, device=dev)
    sh._gather()
    a_sh = sh.accelerations().clone()
    full = torch.empty((n, 2), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(full, a_sh)
    U_sh, dV_sh = sh.potential_and_dVdeps()
    if rank == 0:
        one = LN.LargeNSimulation(m, q, v, softening=1e-3, device=dev, distributed=False)
        a_one = one.accelerations()
        U1, dV1 = one.potential_and_dVdeps()
        err = float((full - a_one).abs().max() / a_one.abs().max())
        print(f"[largeN force n={n}] max rel diff sharded vs single {err:.2e}; U {U_sh:.9e} vs {U1:.9e}", flush=True)
        ok &= err < 1e-6 and abs(U_sh - U1) < 1e-9 * abs(U1) and abs(dV_sh - dV1) < 1e-9 * abs(dV1)
    # (3) large-N ham_soft
    n = 1 << 13
    m, q, v = LN.make_disc(n, seed=6)
    hs = LN.LargeNHamSoftSimulation(m, q, v, softening=0.02, initial_dt=2e-3, device=dev)
    h = 2e-3 / hs.frozen_n_sub
    for _ in range(2):
        hs.strang_step(h)
    qf = hs.xym[:, :2].clone()
    if rank == 0:
        one = LN.LargeNHamSoftSimulation(m, q, v, softening=0.02, initial_dt=2e-3, device=dev, distributed=False)
        for _ in range(2):
            one.strang_step(h)
        err = float((qf - one.xym[:, :2]).abs().max())
        print(f"[largeN ham_soft n={n}] n_sub {hs.frozen_n_sub} vs {one.frozen_n_sub}; eps {hs.eps:.9e} vs {one.eps:.9e}; "
              f"pi {hs.pi:.6e} vs {one.pi:.6e}; max |dq| {err:.2e}", flush=True)
        ok &= hs.frozen_n_sub == one.frozen_n_sub and abs(hs.eps - one.eps) < 1e-7 * abs(one.eps) and err < 1e-6
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) == 1 else 1)


if __name__ == "__main__":
    main()
