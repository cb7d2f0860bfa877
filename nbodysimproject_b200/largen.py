"""Single large-N direct-sum system (BASELINE.json configs[4], "C5").

The reference has no large-N path: gravitational_force materialises (N,N,2) fp64 arrays (geometry_cache.py:30,
16 N^2 bytes).  This module runs the same formulas (forces.py:63-75, 77-112, potential.py:23-64) with O(N) memory:
fp32 pair arithmetic in `nb_largeN_accel_f32`, j-tiles staged through shared memory, fp64 accumulation across
tiles.  Multi-GPU: rank r owns the i-block [r N/P, (r+1) N/P); after every drift the packed (x, y, m, 0) slices
are all-gathered IN PLACE over NCCL/NVLink (the drift kernel writes straight into this rank's slice of the gather
buffer, so there is no separate pack step); scalar sums go through one small all-reduce.
"""
from __future__ import annotations

import math
import os
import time

import numpy as np

from . import _lib as L


class LargeNSimulation:
    def __init__(self, masses, positions, velocities=None, G: float = 1.0, softening: float = 1e-3,
                 integrator_mode: str = "verlet", device=None, group=None):
        torch = L.require_cuda()
        import torch.distributed as dist
        self.torch = torch
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.group = group
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        m = np.asarray(masses, dtype=np.float64)
        q = np.asarray(positions, dtype=np.float64).reshape(-1, 2)
        v = np.zeros_like(q) if velocities is None else np.asarray(velocities, dtype=np.float64).reshape(-1, 2)
        self.n = int(m.size)
        if self.n % self.world != 0:
            raise L.NBodyB200Error("N must be divisible by the number of ranks")
        if integrator_mode not in ("verlet", "yoshida4"):
            raise L.NBodyB200Error("LargeNSimulation supports verlet / yoshida4 kick-drift stepping")
        self.mode = integrator_mode
        self.ni = self.n // self.world
        self.i0 = self.rank * self.ni
        self.G = float(G)
        self.eps = float(softening)
        # COM velocity removal like the facade (simulation.py:85-86)
        v = v - np.sum(m[:, None] * v, axis=0) / np.sum(m)
        xym = np.zeros((self.n, 4), dtype=np.float32)
        xym[:, 0:2] = q
        xym[:, 2] = m
        self.xym = torch.as_tensor(xym).to(self.device)                     # gather buffer (all particles)
        self.vel = torch.as_tensor(v[self.i0:self.i0 + self.ni].astype(np.float32)).to(self.device).contiguous()
        self.acc = torch.zeros((self.ni, 2), dtype=torch.float32, device=self.device)
        self.sums = torch.zeros((2,), dtype=torch.float64, device=self.device)
        self._have_acc = False
        self.force_evals = 0

    # ---- pieces ---------------------------------------------------------------------------------
    @property
    def local(self):
        return self.xym[self.i0:self.i0 + self.ni]

    def accelerations(self, with_sums: bool = False):
        torch = self.torch
        if with_sums:
            self.sums.zero_()
        with torch.cuda.device(self.device):
            L.check(L.load().nb_largeN_accel_f32(L.ptr(self.xym), self.n, self.i0, self.ni, self.eps, self.G,
                                                 L.ptr(self.acc), L.ptr(self.sums) if with_sums else None,
                                                 L.stream_ptr()), "nb_largeN_accel_f32")
        self.force_evals += 1
        self._have_acc = True
        return self.acc

    def _kick_drift(self, kick_h: float, drift_h: float):
        torch = self.torch
        with torch.cuda.device(self.device):
            L.check(L.load().nb_largeN_kick_drift_f32(L.ptr(self.local), L.ptr(self.vel), L.ptr(self.acc), self.ni,
                                                      float(kick_h), float(drift_h), L.stream_ptr()),
                    "nb_largeN_kick_drift_f32")

    def _gather(self):
        if self.dist is not None and self.world > 1:
            self.dist.all_gather_into_tensor(self.xym, self.local, group=self.group)   # in place

    def _verlet_kernel(self, h: float):
        """integration_scheme_base.py:129-149 with FSAL."""
        if not self._have_acc:
            self.accelerations()
        self._kick_drift(0.5 * h, h)
        self._gather()
        self.accelerations()
        self._kick_drift(0.5 * h, 0.0)

    def step(self, dt: float):
        if self.mode == "yoshida4":
            c = 2.0 ** (1.0 / 3.0)
            w1, w2 = 1.0 / (2.0 - c), -c / (2.0 - c)
            for w in (w1, w2, w1):
                self._verlet_kernel(w * dt)
        else:
            self._verlet_kernel(dt)

    # ---- diagnostics (potential.py:23-64, forces.py:77-112, diagnostics.py:63-67) -------------------------
    def potential_and_dVdeps(self):
        """U = -G sum_{i<j} m_i m_j / rho and dV/deps = G eps sum_{i<j} m_i m_j / rho^3 (all ranks)."""
        self.accelerations(with_sums=True)
        s = self.sums.clone()
        if self.dist is not None and self.world > 1:
            self.dist.all_reduce(s, group=self.group)
        U = -self.G * 0.5 * float(s[0])
        dV = self.G * self.eps * 0.5 * float(s[1]) if self.eps != 0.0 else 0.0
        return U, dV

    def kinetic_energy(self):
        m = self.local[:, 2].double()
        T = 0.5 * (m * (self.vel.double() ** 2).sum(1)).sum()
        if self.dist is not None and self.world > 1:
            self.dist.all_reduce(T, group=self.group)
        return float(T)

    def momentum(self):
        m = self.local[:, 2].double()
        P = (m[:, None] * self.vel.double()).sum(0)
        Lz = (m * (self.local[:, 0].double() * self.vel[:, 1].double() - self.local[:, 1].double() * self.vel[:, 0].double())).sum()
        out = self.torch.cat([P, Lz[None]])
        if self.dist is not None and self.world > 1:
            self.dist.all_reduce(out, group=self.group)
        return out.cpu().numpy()


def torch_sum(t):
    return float(t.sum())


def make_disc(n: int, seed: int = 0):
    """SURVEY.md section 8d C5 inputs: positions N(0,1)^2, masses U(0.5,1.5)/N, roughly virial tangential velocities."""
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((n, 2))
    m = rng.uniform(0.5, 1.5, n) / n
    r = np.linalg.norm(q, axis=1)
    menc = 1.0 - np.exp(-0.5 * r * r)                     # enclosed mass of a 2-D Gaussian disc
    vc = np.sqrt(menc / np.maximum(r, 1e-3))
    t = np.stack([-q[:, 1], q[:, 0]], 1) / np.maximum(r, 1e-12)[:, None]
    v = t * vc[:, None] * 0.7 + rng.standard_normal((n, 2)) * 0.1
    return m, q, v


def bench_largen(args, world, rank, local, dev):
    """pair-interactions/s of one force evaluation over all ordered pairs of an N-particle system
    (strong scaling: N fixed, i-blocks sharded, one in-place position all-gather per evaluation)."""
    import torch
    import torch.distributed as dist
    n = int(args.n)
    m, q, v = make_disc(n, seed=1)
    sim = LargeNSimulation(m, q, v, G=1.0, softening=1e-3, device=dev)

    def step():
        sim._gather()
        sim.accelerations()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        tt = torch.tensor([t], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt[0])
    # e2e: host positions in, host accelerations out, every step
    xym_h = sim.xym.cpu().pin_memory()
    acc_h = torch.empty((sim.ni, 2), dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sim.xym.copy_(xym_h, non_blocking=True)
        sim.accelerations()
        acc_h.copy_(sim.acc, non_blocking=True)
        torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    if rank != 0:
        return None
    pairs = float(n) * float(n) * args.steps
    peak32 = L.peak_flops(1, local)
    line = {
        "metric": "pair-interactions/s at N=2^20", "value": pairs / t, "unit": "pair-interactions/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C5 single large-N direct sum, N={n}, Plummer softening 1e-3, one force evaluation "
                               "(+ in-place NCCL position all-gather when sharded) per step",
                   "l2_note": "j-array (16 B x N) is L2-resident by design; inputs are 16 MB at N=2^20"},
        "e2e": {"value": pairs / t_e2e, "unit": "pair-interactions/s", "h2d_bytes_per_step": int(n * 16),
                "d2h_bytes_per_step": int(sim.ni * 8)},
        "gpu_launches": 2 * args.steps,
        "roofline": {"bound": "fp32", "kernel": "largeN_accel_kernel", "achieved": 14.0 * pairs / t * 1e-12 / world,
                     "peak": peak32, "unit": "TFLOP/s", "frac": 14.0 * pairs / t * 1e-12 / world / peak32,
                     "traffic": None, "peak_source": "nb_peak_flops(1): register-resident FFMA micro-benchmark"},
    }
    return line
