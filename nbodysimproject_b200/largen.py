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
                 integrator_mode: str = "verlet", device=None, group=None, distributed: bool = True,
                 spatial_sort: bool = False):
        """spatial_sort: store the particles along a Morton (Z-order) curve.  Results do not depend on the storage
        order; `positions` / `velocities` / `accelerations_in_input_order` give the caller's order back.  The ham_soft
        subclass switches it on: its eps* passes skip tiles whose particles are all farther than 9.35 h apart."""
        torch = L.require_cuda()
        import torch.distributed as dist
        self.torch = torch
        self.dist = dist if (distributed and dist.is_available() and dist.is_initialized()) else None
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
        self.order = None                    # storage index -> input index
        if spatial_sort and self.n > 1:
            self.order = morton_order(q)
            m, q, v = m[self.order], q[self.order], v[self.order]
        xym = np.zeros((self.n, 4), dtype=np.float32)
        xym[:, 0:2] = q
        xym[:, 2] = m
        self.xym = torch.as_tensor(xym).to(self.device)                     # gather buffer (all particles)
        self.vel = torch.as_tensor(v[self.i0:self.i0 + self.ni].astype(np.float32)).to(self.device).contiguous()
        self.acc = torch.zeros((self.ni, 2), dtype=torch.float32, device=self.device)
        self.sums = torch.zeros((2,), dtype=torch.float64, device=self.device)
        self.acc_ws = torch.empty((self.ni, 2), dtype=torch.float64, device=self.device)   # fp64 accumulators of a call
        self.variant = -1                                                                  # kernel variant (A-B tests)
        self._have_acc = False
        self.force_evals = 0

    # ---- pieces ---------------------------------------------------------------------------------
    @property
    def local(self):
        return self.xym[self.i0:self.i0 + self.ni]

    def _to_input_order(self, a):
        """numpy array indexed by storage position (all N particles) -> the caller's particle order"""
        if self.order is None:
            return a
        out = np.empty_like(a)
        out[self.order] = a
        return out

    def _gather_local(self, t):
        if self.dist is not None and self.world > 1:
            out = self.torch.empty((self.n,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
            self.dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
            return out
        return t

    @property
    def positions(self):
        """[N, 2] fp64 positions in the caller's order (all ranks hold all positions)."""
        return self._to_input_order(self.xym[:, :2].double().cpu().numpy())

    @property
    def velocities(self):
        """[N, 2] fp64 velocities in the caller's order (gathered over the ranks)."""
        return self._to_input_order(self._gather_local(self.vel).double().cpu().numpy())

    def accelerations(self, with_sums: bool = False):
        torch = self.torch
        if with_sums:
            self.sums.zero_()
        with torch.cuda.device(self.device):
            L.check(L.load().nb_largeN_accel_f32(L.ptr(self.xym), self.n, self.i0, self.ni, self.eps, self.G,
                                                 L.ptr(self.acc), L.ptr(self.sums) if with_sums else None,
                                                 L.ptr(self.acc_ws), int(self.variant), L.stream_ptr()),
                    "nb_largeN_accel_f32")
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


LN_DENSITY, LN_EPSGRAD, LN_UNITGRAD, LN_TAUMIN = 0, 1, 2, 3
_LOG2E = 1.4426950408889634


class LargeNHamSoftSimulation(LargeNSimulation):
    """ham_soft (adaptive-epsilon Strang split) for ONE large-N system: BASELINE.json configs[4].

    Same flow as the small-N path -- S(h/2) V(h/2) T(h) V(h/2) S(h/2) on (q, p, eps, pi),
    hamsoft_stepper.py:247-308 -- with the reference's O(N^2) Python pair loops replaced by tile-streamed GPU passes
    (`nb_largeN_pass_f32`) and, as SURVEY.md section 7 step 5 prescribes for large N, the ANALYTIC eps* gradient
    (`_production_grad`, hamsoft_eps_model.py:451-556, sign-aligned with softening.py:86-131 exactly as the
    reference's fallback branch does, :200-230) instead of the 4N-solve central difference.  eps, pi and every
    scalar of the spring rotation are fp64 host values replicated on all ranks; reductions go through
    torch.distributed.  Constructor calibration follows hamiltonian_softening_integrator.py:47-141.
    """

    def __init__(self, masses, positions, velocities=None, G: float = 1.0, softening: float = 1e-3,
                 min_softening: float = 0.0, k_soft: float = 1.0e3, k_wall: float = 1.0e9, barrier_exponent: int = 5,
                 theta_cap: float = 0.1, theta_imp: float = 0.5, alpha: float = 0.1, eta: float = 1.35,
                 chi_pi: float = 0.2, j_max_cap: float = 0.02, initial_dt: float = 0.01, use_soft_barrier: bool = True,
                 disable_barrier: bool = False, split_n_max: int = 50, device=None, group=None,
                 distributed: bool = True, spatial_sort: bool = True, cull: bool = True):
        min_softening = max(0.0, float(min_softening))
        softening = float(softening)
        if softening < 0.0:
            softening = min_softening
        if min_softening == 0.0 and softening > 0.0:
            min_softening = 0.1 * softening                         # simulation.py:88-114
        s0 = max(softening, min_softening)
        super().__init__(masses, positions, velocities, G=G, softening=s0, integrator_mode="verlet", device=device,
                         group=group, distributed=distributed, spatial_sort=spatial_sort)
        self.cull = bool(cull)
        torch = self.torch
        self.mode = "ham_soft"
        self.s0 = s0
        self.eps_min, self.eps_max = float(min_softening), 10.0 * s0
        self.pi = 0.0
        self.k_soft, self.mu_soft, self.k_wall, self.n_exp = float(k_soft), 1.0, float(k_wall), int(barrier_exponent)
        self.theta_cap, self.theta_imp, self.alpha_cfg, self.eta = float(theta_cap), float(theta_imp), alpha, float(eta)
        self.chi_pi, self.j_max_cap = float(chi_pi), float(j_max_cap)
        self.soft_policy = bool(use_soft_barrier) and not bool(disable_barrier)
        # hamiltonian_softening_integrator.py:96-108: "reflection" unless the soft barrier is on; disable_barrier switches
        # the folds off as well (hamsoft_barrier_controller.py:46-47)
        self.reflect_policy = (not bool(use_soft_barrier)) and not bool(disable_barrier)
        self.split_n_max = int(split_n_max)
        self.alpha_run = None
        self.n_passes = 0
        self.last_sweeps = 0
        self.m64 = self.local[:, 2].double()
        n_pad = (self.n + 1) // 2 * 2
        self.jaux = torch.zeros((n_pad, 2), dtype=torch.float32, device=self.device)
        self.out64 = torch.zeros((self.ni, 2), dtype=torch.float64, device=self.device)
        self.acc_sums = torch.zeros((2,), dtype=torch.float64, device=self.device)
        n_tiles = (self.n + 511) // 512
        self.boxes = torch.zeros((n_tiles, 8), dtype=torch.float32, device=self.device)   # NB_LN_BOX_FLOATS per NB_LN_TILE
        self._pos_version = 0            # bumped by every drift: tile boxes and the legacy direction depend on q only
        self._boxes_version = (-1, False)
        self._unit_cache = None
        # ---- constructor calibration
        self._calibrate_from_initial_conditions()
        if not (math.isfinite(self.k_soft) and self.k_soft > 0.0):
            em = self.eps_min if (math.isfinite(self.eps_min) and self.eps_min > 0.0) else max(self.s0 * 0.1, 1e-12)
            self.k_soft = 8.0 * self.G * float(self._allsum(self.m64.sum())) ** 2 / em ** 3
        self._calibrate_mu_from_timescales()
        self._freeze_production_schedule(float(initial_dt))

    # ---- collectives on scalars --------------------------------------------------------------------
    def _allsum(self, t):
        if self.dist is not None and self.world > 1:
            t = t.clone()
            self.dist.all_reduce(t, group=self.group)
        return t

    def _allmax(self, t):
        if self.dist is not None and self.world > 1:
            t = t.clone()
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t

    def _allmin(self, t):
        if self.dist is not None and self.world > 1:
            t = t.clone()
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return t

    def _tile_boxes(self, with_jaux: bool):
        """Bounding boxes (+ the smallest exponent scale of jaux) of the 512-particle tiles for the current positions."""
        key = (self._pos_version, bool(with_jaux))
        if with_jaux or self._boxes_version != key:
            torch = self.torch
            with torch.cuda.device(self.device):
                L.check(L.load().nb_largeN_tile_boxes_f32(L.ptr(self.xym), L.ptr(self.jaux) if with_jaux else None, self.n,
                                                          L.ptr(self.boxes), L.stream_ptr()), "nb_largeN_tile_boxes_f32")
            self._boxes_version = key
        return self.boxes

    def _pass(self, kind, iparam=None, eps=0.0):
        torch = self.torch
        boxes = None
        if self.cull and kind in (LN_DENSITY, LN_EPSGRAD):
            boxes = self._tile_boxes(kind == LN_EPSGRAD)
        with torch.cuda.device(self.device):
            L.check(L.load().nb_largeN_pass_f32(kind, L.ptr(self.xym), L.ptr(self.jaux), self.n, self.i0, self.ni,
                                                L.ptr(iparam), float(eps), L.ptr(self.out64), L.ptr(boxes),
                                                L.stream_ptr()), "nb_largeN_pass_f32")
        self.n_passes += 1
        if boxes is None:
            self.n_full_passes = getattr(self, "n_full_passes", 0) + 1
        return self.out64

    # ---- reflection policy: fold epsilon into [eps_min, eps_max], flipping pi (hamsoft_barrier_controller.py:27-69,
    #      hamsoft_utils.py:105-184 reflect_if_needed) -- eps, pi are replicated host scalars
    def _fold(self):
        if not self.reflect_policy:
            return
        a, b = float(self.eps_min), float(self.eps_max)
        R = b - a
        if not math.isfinite(R) or R <= 0.0:
            self.eps, self.pi = a, -float(self.pi)
            return
        y = (float(self.eps) - a) % (2.0 * R)           # Python float modulo: sign of the divisor
        if y <= R:
            self.eps = a + y
        else:
            self.eps, self.pi = b - (y - R), -float(self.pi)

    # ---- barrier (barrier.py:66-113) ----------------------------------------------------------------
    def _fbar(self, eps):
        if not self.soft_policy or not (math.isfinite(self.k_wall) and self.k_wall > 0.0):
            return 0.0
        e = max(2, self.n_exp) - 2
        la, rb = max(0.0, self.eps_min - eps), max(0.0, eps - self.eps_max)
        left = (1.0 if e == 0 else la ** e) if la > 0.0 else 0.0
        right = (1.0 if e == 0 else rb ** e) if rb > 0.0 else 0.0
        return self.k_wall * (left - right)

    def _alpha(self):
        if self.alpha_run is not None and self.alpha_run > 0.0:
            return float(self.alpha_run)
        if isinstance(self.alpha_cfg, (int, float)) and self.alpha_cfg > 0.0:
            return float(self.alpha_cfg)
        return 1.0

    # ---- eps* model -----------------------------------------------------------------------------------
    def solve_hi(self):
        """hamsoft_eps_model.py:316-400: Jacobi sweeps (<= 8, tol 1e-6) from the CURRENT epsilon; one DENSITY pass each."""
        torch = self.torch
        lo, hi = self.eps_min, self.eps_max
        if hi < lo:
            lo, hi = hi, lo
        floor = max(lo, 1.0e-12)
        cap = max(floor, hi)
        h0 = float(self.eps)
        if not math.isfinite(h0) or h0 <= 0.0:
            h0 = 1.0
        h0 = min(max(h0, floor), cap)
        h = torch.full((self.ni,), h0, dtype=torch.float64, device=self.device)
        it = 0
        while it < 8:
            S = self._pass(LN_DENSITY, h.float().contiguous())
            Sigma = torch.clamp_min(S[:, 0] / (math.pi * h * h), 1.0e-30)
            hn = self.eta * torch.sqrt(self.m64 / Sigma)
            hn = torch.where(torch.isfinite(hn) & (hn > 0.0), hn, h).clamp(floor, cap)
            changed = float(self._allmax(((hn - h).abs() / h.clamp_min(1.0e-12)).max()))
            h = hn
            if changed < 1.0e-6:
                break
            it += 1
        self.last_sweeps = min(it + 1, 8)
        return h

    def _softmin(self, h):
        a = self._alpha()
        t = -h / a
        tmax = float(self._allmax(t.max()))
        ex = self.torch.exp(t - tmax)
        den = float(self._allsum(ex.sum()))
        return a, tmax, ex, den

    def eps_target(self, h=None):
        """hamsoft_eps_model.py:240-289 eps_target_production: eps* = -alpha ln sum_i exp(-h_i/alpha), clamped."""
        if h is None:
            h = self.solve_hi()
        a, tmax, _, den = self._softmin(h)
        es = self.s0 if (den <= 0.0 or not math.isfinite(den)) else -a * (tmax + math.log(den))
        if self.soft_policy:
            lo, hi = min(self.eps_min, self.eps_max), max(self.eps_min, self.eps_max)
            es = min(max(es, lo), hi)
        return float(es)

    def production_grad(self, h):
        """hamsoft_eps_model.py:451-556 in gather form + the sign alignment of :200-230 (softening.py:86-131)."""
        torch = self.torch
        a, tmax, ex, den = self._softmin(h)
        if den <= 0.0 or not math.isfinite(den):
            return torch.zeros((self.ni, 2), dtype=torch.float64, device=self.device)
        floor = max(self.eps_min, 1.0e-12)
        hj = h.clamp_min(max(1.0e-12, 0.1 * floor))
        w = ex / den
        S = self._pass(LN_DENSITY, hj.float().contiguous())
        c = 1.0 / (math.pi * hj * hj)
        Sigma = torch.clamp_min(c * S[:, 0], 1.0e-30)
        Sd = c * (-2.0 / hj * S[:, 0] + 2.0 / (hj * hj * hj) * S[:, 1])
        Om = 1.0 + hj * Sd / (2.0 * Sigma)
        Om = torch.where(torch.isfinite(Om) & (Om != 0.0), Om, torch.ones_like(Om))
        Pi = -hj / (2.0 * Sigma * Om)
        s_i = -w * Pi
        A = s_i * (-2.0 / (math.pi * hj ** 4))
        loc = torch.stack([-_LOG2E / (hj * hj), A], 1).float()
        self.jaux[self.i0:self.i0 + self.ni] = loc
        if self.dist is not None and self.world > 1:
            self.dist.all_gather_into_tensor(self.jaux[:self.n], self.jaux[self.i0:self.i0 + self.ni], group=self.group)
        g = self._pass(LN_EPSGRAD).clone()
        g = torch.where(torch.isfinite(g), g, torch.zeros_like(g))
        # legacy gradient = -c_pref * u with c_pref > 0; u depends on the positions only, and the S half-flow that ends a
        # sub-step and the one that opens the next see the same positions: one long-range pass serves both
        if self._unit_cache is None or self._unit_cache[0] != self._pos_version:
            self._unit_cache = (self._pos_version, self._pass(LN_UNITGRAD).clone())
        u = self._unit_cache[1]
        dot = float(self._allsum(-(g * u).sum()))
        if math.isfinite(dot) and dot < 0.0:
            g = -g
        return g

    def eps_star_and_grad(self):
        h = self.solve_hi()
        return self.eps_target(h), self.production_grad(h)

    # ---- calibration and schedule -----------------------------------------------------------------------
    def _gather_vector(self, t):
        if self.dist is not None and self.world > 1:
            out = self.torch.empty((self.n,), dtype=t.dtype, device=self.device)
            self.dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
            return out
        return t

    def _calibrate_from_initial_conditions(self):
        """hamsoft_eps_model.py:645-729."""
        a_seed = float(self.alpha_cfg) if isinstance(self.alpha_cfg, (int, float)) and self.alpha_cfg > 0 else max(self.eps, 1e-12)
        h0 = self._gather_vector(self.solve_hi())
        hs = self.torch.sort(h0).values                     # np.median: mean of the two middle values for even n
        med = float(0.5 * (hs[(self.n - 1) // 2] + hs[self.n // 2]))
        if not math.isfinite(med) or med <= 0.0:
            med = a_seed
        self.alpha_run = 0.3 * med
        if not math.isfinite(self.alpha_run) or self.alpha_run <= 0.0:
            self.alpha_run = a_seed
        cand = 0.25 * med
        emin0 = self.eps_min if (math.isfinite(self.eps_min) and self.eps_min >= 0.0) else 0.0
        emax = self.eps_max if (math.isfinite(self.eps_max) and self.eps_max > 0.0) else 10.0 * self.s0
        if not math.isfinite(cand):
            cand = emin0
        cand = min(cand, emax)
        new = emin0 if emin0 >= cand else cand
        self.eps_min = float(min(new, emax))
        if self.eps < self.eps_min:
            self.eps = float(self.eps_min)

    def _tau_grav_soft(self, fallback):
        """hamiltonian_softening_integrator.py:251-296: min over pairs of sqrt(rho^3 / (G (m_i + m_j)))."""
        tau = math.inf
        if self.n >= 2 and self.G != 0.0:
            out = self._pass(LN_TAUMIN, eps=self.eps)
            vmin = float(self._allmin(out.view(-1)[:self.ni].min()))
            if vmin < 1.0e38 and vmin > 0.0:
                tau = math.sqrt(vmin / self.G)
        if (not math.isfinite(tau)) or tau <= 0.0:
            tau = fallback
        return float(tau)

    def _calibrate_mu_from_timescales(self):
        tau = self._tau_grav_soft(1.0)
        k = self.k_soft if (math.isfinite(self.k_soft) and self.k_soft > 0.0) else 0.0
        om = 8.0 / tau if tau > 0.0 else 0.0
        mu = (k / (om * om) if k > 0.0 else 1.0) if om > 0.0 else 1.0
        if (not math.isfinite(mu)) or mu <= 0.0:
            mu = 1.0
        self.mu_soft, self.omega_spr0 = float(mu), float(om)

    def dV_d_epsilon(self):
        """forces.py:77-112 through the force kernel's fused scalar sums."""
        _, dV = self.potential_and_dVdeps()
        return dV

    def _estimate_pi_budget_h(self, dt_abs):
        """hamiltonian_softening_integrator.py:1125-1221."""
        k = self.k_soft
        if (not math.isfinite(k)) or k <= 0.0:
            return float(dt_abs)
        Delta = self.eps - self.eps_target()
        s0 = self.s0 if (math.isfinite(self.s0) and self.s0 > 0.0) else 1.0
        d_eff = max(abs(Delta), 1.0e-4 * s0)
        dV = self.dV_d_epsilon() if (self.n >= 2 and self.G != 0.0) else 0.0
        dB = -self._fbar(self.eps) if self.soft_policy else 0.0
        tot = max(abs(dV + dB), 1.0e-16)
        h_pi = (2.0 * self.chi_pi * math.sqrt(k) * d_eff) / tot
        if (not math.isfinite(h_pi)) or h_pi < 0.0:
            h_pi = float(dt_abs)
        return float(h_pi)

    def _freeze_production_schedule(self, dt_user):
        """hamiltonian_softening_integrator.py:986-1119."""
        dt_abs = abs(float(dt_user))
        if (not math.isfinite(dt_abs)) or dt_abs <= 0.0:
            dt_abs = 1.0e-2
        tau = self._tau_grav_soft(dt_abs)
        om = self.omega_spr0
        if (not math.isfinite(om)) or om <= 0.0:
            om = 8.0 / tau if tau > 0.0 else 0.0
            self.omega_spr0 = om
        theta_cap = self.theta_cap if (math.isfinite(self.theta_cap) and self.theta_cap > 0.0) else 0.1
        h_g = 0.9 * tau
        h_o = theta_cap / om if om > 0.0 else math.inf
        h_theta = min(h_g, h_o) if (math.isfinite(h_o) and h_o > 0.0) else h_g
        h_pi = self._estimate_pi_budget_h(dt_abs)
        if (not math.isfinite(h_pi)) or h_pi <= 0.0:
            h_pi = dt_abs
        h_sub = min(h_theta, h_pi)
        if (not math.isfinite(h_sub)) or h_sub <= 0.0:
            h_sub = dt_abs
        n_sub = int(math.ceil(dt_abs / h_sub)) if h_sub > 0.0 else 1
        self.frozen_n_sub = max(1, n_sub)
        self.macro_dt_frozen = dt_abs
        self.h_theta, self.h_pi = h_theta, h_pi

    def strang_substeps(self, dt):
        """hamiltonian_softening_integrator.py:781-888 (mu floor :232-242, frozen n_sub within 1 % of the frozen dt)."""
        dt_abs = abs(float(dt))
        mu_macro = self.k_soft * (dt_abs / self.theta_imp) ** 2
        if math.isfinite(self.k_soft) and self.k_soft > 0.0 and self.mu_soft < mu_macro:
            self.mu_soft = float(mu_macro)
        prev = self.macro_dt_frozen
        if not (prev > 0.0 and abs(dt_abs - prev) / prev <= 0.01):
            self._freeze_production_schedule(dt_abs)
        return int(self.frozen_n_sub)

    # ---- flows (hamsoft_stepper.py:47-308, hamsoft_flows.py:427-762, 1102-1132) ---------------------------------
    def s_half(self, h):
        torch = self.torch
        dt = 0.5 * float(h)
        self._fold()                                            # hamsoft_stepper.py:107-113
        eps0, pi0 = float(self.eps), float(self.pi)
        es, grad = self.eps_star_and_grad()
        k, mu = self.k_soft, self.mu_soft
        om = math.sqrt(k / mu) if (k > 0.0 and mu > 0.0) else 0.0
        th = om * dt
        if abs(th) < 1.0e-8:
            th2 = th * th
            sn = th - th2 * th / 6.0 + th2 * th2 * th / 120.0
            cs = 1.0 - th2 / 2.0 + th2 * th2 / 24.0
        else:
            sn, cs = math.sin(th), math.cos(th)
        kick1 = -0.5 * dt * (-self._fbar(eps0)) if self.soft_policy else 0.0
        D0 = eps0 - es
        pin = pi0 + kick1
        if om != 0.0 and mu != 0.0:
            mo = math.sqrt(mu * max(k, 0.0))
            dlt = D0 * cs + (pin / (mu * om)) * sn
            eta_t = pin * cs - mo * D0 * sn
            den = mu * om * om
            I = (D0 / om) * sn + (pin / den) * (1.0 - cs) if den != 0.0 else 0.0
        else:
            dlt, eta_t, I = D0, pin, 0.0
        eps_rot = es + dlt
        kick2 = -0.5 * dt * (-self._fbar(eps_rot)) if self.soft_policy else 0.0
        J = k * I
        v64 = self.vel.double()
        pn = float(self._allmax((self.m64 * torch.sqrt((v64 * v64).sum(1))).max()))
        gn = float(self._allmax(torch.sqrt((grad * grad).sum(1)).max()))
        p_scale = max(pn, 1.0e-12)
        dp_inf = abs(J) * gn
        thr = self.j_max_cap * p_scale
        Ja = J * (thr / dp_inf) if (dp_inf > thr and dp_inf > 0.0) else J
        self.vel += (Ja * grad / self.m64[:, None]).float()
        self.eps = float(eps_rot)
        self.pi = float(eta_t + kick2)
        self._fold()                                            # hamsoft_stepper.py:72-80
        self.taps = dict(eps_star=es, J=J, J_applied=Ja, theta=th, sweeps=self.last_sweeps)

    def v_half_kick(self, h):
        hh = 0.5 * float(h)
        e = float(self.eps)
        dU = 0.0
        if self.n >= 2 and self.G != 0.0:
            self.accelerations(with_sums=True)               # force kernel at the CURRENT epsilon + fused sum m_i m_j / rho^3
            s = self._allsum(self.sums)
            dU = self.G * e * 0.5 * float(s[1]) if e != 0.0 else 0.0
            self._kick_drift(hh, 0.0)
        dB = -self._fbar(e) if self.soft_policy else 0.0
        self.pi = float(self.pi) - (dU + dB) * hh

    def strang_step(self, h):
        self._fold()                                            # hamsoft_stepper.py:261-264
        self.s_half(h)
        self.v_half_kick(h)
        self._kick_drift(0.0, h)
        self._pos_version += 1
        self._gather()
        self.v_half_kick(h)
        self.s_half(h)
        self._fold()                                            # hamsoft_stepper.py:300-303

    def step(self, dt):
        """hamiltonian_softening_integrator.py:496-557."""
        if dt == 0.0 or self.n == 0:
            return
        n_pred = max(1, self.strang_substeps(dt))
        h = float(dt) / float(n_pred)
        for _ in range(n_pred):
            self.strang_step(h)
        self.n_sub_last = n_pred

    def extended_hamiltonian(self):
        """hamsoft_energy.py:48-162: T + U + pi^2/2mu + k/2 (eps - eps*)^2 + S_bar."""
        U, _ = self.potential_and_dVdeps()
        T = self.kinetic_energy()
        es = self.eps_target()
        H = T + U + self.pi * self.pi / (2.0 * self.mu_soft) + 0.5 * self.k_soft * (self.eps - es) ** 2
        if self.soft_policy and math.isfinite(self.k_wall) and self.k_wall > 0.0 and self.n_exp >= 2:
            a, b = min(self.eps_min, self.eps_max), max(self.eps_min, self.eps_max)
            p = self.n_exp - 1
            H += (self.k_wall / p) * (max(0.0, a - self.eps) ** p + max(0.0, self.eps - b) ** p)
        return H


def morton_order(q, bits: int = 16):
    """Permutation that sorts 2-D points along a Z-order curve (interleaved bits of the quantised coordinates)."""
    q = np.asarray(q, dtype=np.float64)
    lo, hi = q.min(0), q.max(0)
    span = np.maximum(hi - lo, 1e-300)
    g = np.minimum(((q - lo) / span * (1 << bits)).astype(np.uint64), (1 << bits) - 1)

    def spread(x):
        x = x & np.uint64(0xFFFF)
        x = (x | (x << np.uint64(8))) & np.uint64(0x00FF00FF)
        x = (x | (x << np.uint64(4))) & np.uint64(0x0F0F0F0F)
        x = (x | (x << np.uint64(2))) & np.uint64(0x33333333)
        x = (x | (x << np.uint64(1))) & np.uint64(0x55555555)
        return x

    key = spread(g[:, 0]) | (spread(g[:, 1]) << np.uint64(1))
    return np.argsort(key, kind="stable")


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


def measure_force(sim, steps: int, warmup: int = 3):
    """CUDA-event time of `steps` x (in-place position all-gather + one force evaluation); returns seconds (this rank)."""
    torch = sim.torch
    for _ in range(max(warmup, 3)):
        sim._gather()
        sim.accelerations()
    torch.cuda.synchronize()
    if sim.dist is not None and sim.world > 1:
        sim.dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sim._gather()
        sim.accelerations()
    e1.record()
    torch.cuda.synchronize()
    if sim.dist is not None and sim.world > 1:
        sim.dist.barrier()
    t = e0.elapsed_time(e1) * 1e-3
    if sim.dist is not None and sim.world > 1:
        tt = torch.tensor([t], dtype=torch.float64, device=sim.device)
        sim.dist.all_reduce(tt, op=sim.dist.ReduceOp.MAX)
        t = float(tt[0])
    return t


def bench_largen(args, world, rank, local, dev, sampler=None):
    """pair-interactions/s of one force evaluation over all ordered pairs of an N-particle system
    (strong scaling: N fixed, i-blocks sharded, one in-place position all-gather per evaluation), plus the wall
    time of one full ham_soft Strang sub-step S V T V S (adaptive epsilon) on the same particles."""
    import torch
    import torch.distributed as dist
    n = int(args.n)
    m, q, v = make_disc(n, seed=1)
    sim = LargeNSimulation(m, q, v, G=1.0, softening=1e-3, device=dev)
    mark0 = sampler.mark() if sampler is not None else 0
    t = measure_force(sim, args.steps, args.warmup)
    mark1 = sampler.mark() if sampler is not None else 0
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
    if world > 1:
        tt = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt[0])
    # full Strang sub-step of the adaptive-epsilon flow (C5 as BASELINE.json words it)
    strang = None
    n_hs = int(getattr(args, "n_hamsoft", 0) or n)            # C5 as BASELINE.json words it: the same N = 2^20 particles
    if n_hs > 0:
        mh, qh, vh = make_disc(n_hs, seed=1)
        hs = LargeNHamSoftSimulation(mh, qh, vh, softening=2.0 / math.sqrt(n_hs), initial_dt=1e-3, device=dev)
        hsub = 1e-3 / hs.frozen_n_sub
        hs.strang_step(hsub)
        torch.cuda.synchronize()
        p0, f0 = hs.n_passes, hs.force_evals
        t0 = time.perf_counter()
        hs.strang_step(hsub)
        torch.cuda.synchronize()
        ts = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([ts], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ts = float(tt[0])
        n_pass = (hs.n_passes - p0) + (hs.force_evals - f0)
        strang = {"n": n_hs, "ms": 1e3 * ts, "n2_passes": n_pass, "solver_sweeps": hs.last_sweeps,
                  "pair_evaluations_per_s": n_pass * float(n_hs) * n_hs / ts, "frozen_n_sub": hs.frozen_n_sub,
                  "eps": hs.eps, "eps_min": hs.eps_min, "eps_max": hs.eps_max}
    if rank != 0:
        return None
    pairs = float(n) * float(n) * args.steps
    peak32 = L.peak_flops(1, local)
    cpu = None          # filled in by bench.py (the only place allowed to time the oracle)
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
        "roofline": {"bound": "fp32", "kernel": "largeN_accel_x2_kernel", "achieved": 14.0 * pairs / t * 1e-12 / world,
                     "peak": peak32, "unit": "TFLOP/s", "frac": 14.0 * pairs / t * 1e-12 / world / peak32,
                     "traffic": None, "flops_per_pair": 14,
                     "peak_source": "nb_peak_flops(1): register-resident FFMA micro-benchmark, same GPU, same run"},
        "hamsoft_strang_substep": strang, "cpu_baseline": cpu,
        "clocks": sampler.stop(mark0, mark1) if sampler is not None else None,
    }
    return line
