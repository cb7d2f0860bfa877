"""CPU oracle for the NBodySimProject hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference's algorithm for the path
named in BASELINE.json (softened pair kernels -> kick/drift integrators ->
tangent map / MEGNO -> energy diagnostics -> feature row).  It exists so the
CUDA path can be checked on a box that does not have /root/reference.

* Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
  package ``nbodysimproject_b200`` never does; it fails loudly when the CUDA
  library is missing.
* Parity status: PINNED.  The reference ships no tests or golden vectors
  (SURVEY.md section 4), so the oracle is pinned against outputs of the live reference
  itself, generated in the build container by ``oracle/make_golden.py`` and
  committed under ``tests/golden/`` (checked by ``tests/test_oracle_golden.py``).
* Every function cites the reference file:line it restates (paths relative to
  the reference's ``minbody/`` package).

The code is written from the formulas -- it is deliberately a
per-system, loop-over-steps implementation like the reference, so that timing
it on host cores is a fair "reference CPU path" number.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

HP = getattr(np, "float128", np.longdouble)

# --------------------------------------------------------------------------
# L0 pair kernels
# --------------------------------------------------------------------------


def geometry_buffers(pos, eps=0.0):
    """geometry_cache.py:24-39 -- diff_ij = q_i - q_j, r2, (r2+eps^2)^-1.5 with zero diagonal."""
    pos = np.asarray(pos, dtype=float)
    diff = pos[:, None, :] - pos[None, :, :]
    r2 = np.einsum("ijk,ijk->ij", diff, diff)
    inv_r3 = np.zeros_like(r2)
    soft = r2 + eps * eps
    mask = soft > 0.0
    inv_r3[mask] = np.power(soft[mask], -1.5)
    np.fill_diagonal(inv_r3, 0.0)
    return diff, r2, inv_r3


def gravitational_force(q, m, eps=0.0, G=1.0):
    """forces.py:63-75 -- F_i = sum_j -G m_i m_j (r^2+eps^2)^-1.5 (q_i - q_j)."""
    q = np.asarray(q, dtype=float)
    m = np.asarray(m, dtype=float)
    if q.shape[0] < 2 or G == 0.0:
        return np.zeros_like(q)
    diff, _, inv_r3 = geometry_buffers(q, eps)
    pair = -(G * m[:, None] * m[None, :])[..., None] * inv_r3[..., None] * diff
    return pair.sum(axis=1)


def accelerations(q, m, eps, G=1.0):
    """simulation.py:539-581 -- a_i = F_i / m_i."""
    F = gravitational_force(q, m, eps, G)
    return F / np.asarray(m, dtype=float)[:, None]


def dV_d_epsilon(q, m, eps, G=1.0):
    """forces.py:77-112 -- G eps sum_{i<j} m_i m_j (r^2+eps^2)^-1.5 ; 0 if eps == 0."""
    q = np.asarray(q, dtype=float)
    m = np.asarray(m, dtype=float)
    n = q.shape[0]
    if q.ndim != 2 or q.shape[1] != 2 or m.size != n or n < 2 or float(G) == 0.0:
        return 0.0
    e = float(eps)
    if e == 0.0:
        return 0.0
    diff = q[:, None, :] - q[None, :, :]
    r2 = np.sum(diff * diff, axis=-1) + e * e
    iu = np.triu_indices(n, 1)
    r32 = np.power(r2[iu], 1.5)
    return float(G * e * float(np.sum((m[iu[0]] * m[iu[1]]) / r32)))


def softened_potential(q, m, G, eps):
    """potential.py:23-64 -- U = -G sum_{i<j} m_i m_j / sqrt(r^2+eps^2)."""
    q = np.asarray(q, dtype=float)
    m = np.asarray(m, dtype=float).ravel()
    if q.ndim != 2 or q.shape[1] != 2:
        return 0.0
    n = q.shape[0]
    if n < 2 or float(G) == 0.0 or m.size != n:
        return 0.0
    diff = q[:, None, :] - q[None, :, :]
    r2 = np.einsum("ijk,ijk->ij", diff, diff) + float(eps) * float(eps)
    iu = np.triu_indices(n, 1)
    rs = np.sqrt(r2[iu])
    inv = np.zeros_like(rs)
    ok = rs > 0.0
    inv[ok] = 1.0 / rs[ok]
    return float(-float(G) * float(np.sum(m[iu[0]] * m[iu[1]] * inv)))


def variational_accel(q, m, s2, delta_r, G=1.0):
    """tangent_map.py:21-59 -- da_i = G sum_j m_j [d_ij rho^-3 - 3 (D_ij . d_ij) rho^-5 D_ij],
    D_ij = q_j - q_i, d_ij = dr_j - dr_i, rho^2 = r^2 + s2."""
    q = np.asarray(q, dtype=float)
    m = np.asarray(m, dtype=float)
    delta_r = np.asarray(delta_r, dtype=float)
    n = q.shape[0]
    if n < 2 or G == 0.0:
        return np.zeros_like(delta_r)
    diff = q[None, :, :] - q[:, None, :]
    r2 = np.einsum("ijk,ijk->ij", diff, diff) + s2
    np.fill_diagonal(r2, np.inf)
    inv_r2 = 1.0 / r2
    inv_r3 = inv_r2 * np.sqrt(inv_r2)
    d_diff = delta_r[None, :, :] - delta_r[:, None, :]
    dot = np.einsum("ijk,ijk->ij", diff, d_diff)
    coeff = 3.0 * dot * inv_r2 * inv_r3
    term = d_diff * inv_r3[..., None] - coeff[..., None] * diff
    return G * np.sum(m[None, :, None] * term, axis=1)


def remove_center_of_mass_velocity(m, v):
    """physics_utils.py:16-26."""
    m = np.asarray(m, dtype=float)
    v = np.asarray(v, dtype=float)
    if len(m) == 1:
        return v.copy()
    M = float(np.sum(m))
    if M == 0 or v.size == 0:
        return v.copy()
    return v - np.sum(m[:, None] * v, axis=0) / M


# --------------------------------------------------------------------------
# Kepler solver (bug-compatible)  kepler_solver.py:25-107
# --------------------------------------------------------------------------


def kepler_cfunc(z):
    """kepler_solver.py:25-46 -- Stumpff series + the reference's doubling recurrence."""
    z = float(z)
    n = 0
    while abs(z) > 0.1:
        z *= 0.25
        n += 1
    z2 = z * z
    c0 = 1 - z * 0.5 + z2 / 24 - z * z2 / 720 + z2 * z2 / 40320
    c1 = 1 - z / 6 + z2 / 120 - z * z2 / 5040 + z2 * z2 / 362880
    c2 = 0.5 - z / 24 + z2 / 720 - z * z2 / 40320
    c3 = 1 / 6 - z / 120 + z2 / 5040 - z * z2 / 362880
    while n:
        z *= 4
        n -= 1
        c3o, c1o, c2o = c3, c1, c2
        c0 = 1 - z * c2o
        c1 = 1 - z * c3o
        c2 = 0.5 - z * (c3o * (1 + c1o)) * 0.125
        c3 = (c1o - 1) / z
    return c0, c1, c2, c3


def kepler_propagate(r, v, mu, dt, return_iters=False):
    """kepler_solver.py:48-91 -- universal-variable Newton solve, reference semantics
    (uses c1,c2 in f and the 2-cycle exit)."""
    r = np.asarray(r, dtype=float)
    v = np.asarray(v, dtype=float)
    mu = float(mu)
    dt = float(dt)
    r0 = float(math.hypot(r[0], r[1]))
    if r0 < 1e-14:
        out = (r + v * dt, v)
        return out + (0,) if return_iters else out
    vr0 = float(np.dot(r, v) / r0)
    v2 = float(np.dot(v, v))
    alpha = 2 / r0 - v2 / mu
    sqrt_mu = math.sqrt(mu)
    if abs(alpha) > 1e-12:
        chi = sqrt_mu * abs(alpha) * dt
    else:
        chi = sqrt_mu * dt / r0
    prev1 = math.nan
    prev2 = math.nan
    iters = 0
    for _ in range(64):
        iters += 1
        z = alpha * chi * chi
        c0, c1, c2, c3 = kepler_cfunc(z)
        f = r0 * vr0 / sqrt_mu * chi * chi * c1 + (1 - alpha * r0) * chi * chi * chi * c2 + r0 * chi - sqrt_mu * dt
        fp = r0 * vr0 / sqrt_mu * chi * (1 - alpha * chi * chi * c2) + (1 - alpha * r0) * chi * chi * c1 + r0
        if fp == 0:
            break
        chi_new = chi - f / fp
        prev2 = prev1
        prev1 = chi_new
        if chi_new == chi or chi_new == prev2:
            chi = chi_new
            break
        chi = chi_new
    z = alpha * chi * chi
    c0, c1, c2, c3 = kepler_cfunc(z)
    f = 1 - chi * chi * c2 / r0
    g = dt - chi * chi * chi * c3 / sqrt_mu
    r_vec = f * r + g * v
    rn = float(math.hypot(r_vec[0], r_vec[1]))
    if rn == 0:
        out = (r_vec, v)
        return out + (iters,) if return_iters else out
    fdot = sqrt_mu / (rn * r0) * (alpha * chi * chi * c3 - chi)
    gdot = 1 - chi * chi * c2 / rn
    v_vec = fdot * r + gdot * v
    out = (r_vec, v_vec)
    return out + (iters,) if return_iters else out


# --------------------------------------------------------------------------
# classic sub-step schedule  timestep_manager.py:139-253 (k_soft = 0, pi = 0 branch)
# --------------------------------------------------------------------------


def classic_h_sub_ref(q, m, G, dt_user, split_cap=50):
    """timestep_manager.py:139-253 for the classic integrators: tau_spr, tau_eps, tau_imp are
    +inf (Integrator.k_soft = 0, pi = 0), so h = 0.9 * min_ij sqrt(r_ij^3 / (G (m_i+m_j)))
    from UNSOFTENED separations, then capped so that ceil(dt/h) <= split_cap."""
    q = np.asarray(q, dtype=float)
    m = np.asarray(m, dtype=float)
    dt_user = abs(float(dt_user))
    n = q.shape[0]
    if n < 2 or G == 0.0:
        tau = math.inf
    else:
        diff = q[:, None, :] - q[None, :, :]
        r2 = np.einsum("ijk,ijk->ij", diff, diff)
        np.fill_diagonal(r2, np.inf)
        r = np.sqrt(r2)
        r3 = r * r * r
        denom = float(G) * (m[:, None] + m[None, :])
        valid = np.isfinite(r3) & np.isfinite(denom) & (denom > 0.0)
        tau_ij = np.full_like(r3, np.inf)
        tau_ij[valid] = np.sqrt(r3[valid] / denom[valid])
        tau = float(np.min(tau_ij)) if np.any(np.isfinite(tau_ij)) else math.inf
    h = 0.9 * tau
    if not math.isfinite(h) or h <= 0.0:
        h = dt_user if dt_user > 0.0 else 1.0
    if split_cap > 0:
        if math.ceil(dt_user / max(h, 1e-30)) > split_cap:
            h = dt_user / split_cap
    return float(h)


_CBRT2 = 2.0 ** (1.0 / 3.0)
YOSHIDA_W1 = 1.0 / (2.0 - _CBRT2)
YOSHIDA_W2 = -_CBRT2 / (2.0 - _CBRT2)


# --------------------------------------------------------------------------
# classic simulation (verlet / yoshida4 / whfast)   simulation.py, integrator.py
# --------------------------------------------------------------------------


def _barrier_energy(eps, a, b, k_wall=1.0e9, n=5):
    """barrier.py:35-63."""
    if not (np.isfinite(k_wall) and k_wall > 0.0 and n >= 2):
        return 0.0
    if b < a:
        a, b = b, a
    p = n - 1
    return float((k_wall / float(p)) * (max(0.0, a - eps) ** p + max(0.0, eps - b) ** p))


class OracleSim:
    """Restates NBodySimulation for integrator_mode in {verlet, yoshida4, whfast}
    (simulation.py:39-162 ctor, :667-676 step, :319-326 commit_state/snapshot kick)."""

    def __init__(self, masses, positions, velocities=None, G=1.0, softening=1e-3, min_softening=0.0,
                 integrator_mode="verlet", skip_init_corrector=False, skip_cm_recenter=False,
                 initial_dt=0.01, split_n_max=50, corrector_order=5, adaptive_softening=False, adaptive_timestep=None,
                 softening_scale=1.0, k_wall=1.0e9, barrier_exponent=5):
        m = np.asarray(list(masses), dtype=np.float64)
        q = np.asarray(list(positions), dtype=np.float64).reshape(-1, 2).copy()
        if velocities is None or len(velocities) == 0:
            v = np.zeros_like(q)
        else:
            v = np.asarray(list(velocities), dtype=np.float64).reshape(-1, 2).copy()
            if v.shape[0] == 1 and m.size > 1:
                v = np.repeat(v, m.size, axis=0)
        self.m, self.q, self.v = m, q, v
        self.n = int(m.size)
        if not skip_cm_recenter:                                    # simulation.py:85-86
            self.v = remove_center_of_mass_velocity(self.m, self.v)
        min_softening = max(0.0, min_softening)                     # :88-94
        if softening < 0.0:
            softening = min_softening
        if min_softening == 0.0 and softening > 0.0:
            min_softening = 0.1 * softening
        self.min_softening = float(min_softening)
        self.G = float(G)
        mode = str(integrator_mode)
        # simulation.py:62-74: adaptive softening implies the adaptive-timestep flag (whose only live effect is
        # that the constructor corrector is skipped, :150-157)
        self.adaptive_softening = bool(adaptive_softening)
        self.adaptive_timestep = bool(adaptive_timestep) if adaptive_timestep is not None else False
        if self.adaptive_softening:
            self.adaptive_timestep = True
        self.softening_scale = float(softening_scale)
        self.k_wall, self.barrier_exponent = float(k_wall), int(barrier_exponent)
        self.softening_energy_delta = 0.0
        self.pending_energy_delta = 0.0
        if self.G == 0.0 and mode != "ham_soft":                    # :101-102
            mode = "verlet"
        if mode == "whfast" and self.n > 0:                         # :103-111
            if self.adaptive_softening:
                mode = "verlet"
            elif np.max(self.m) / np.sum(self.m) < 0.2:
                mode = "verlet"
        self.s0 = float(max(softening, self.min_softening))         # softening_manager.py:48
        self.s = self.s0
        self.eps_attr = self.s0        # sim._epsilon: set once in the constructor (simulation.py:116), never by the classic refresh
        self.step_s2 = self.s * self.s
        self.max_softening = 10.0 * self.s0
        if self.s > 0.0 and mode == "whfast":                       # :119-120
            mode = "verlet"
        self.mode = mode
        self.split_n_max = int(split_n_max)
        self.initial_dt = float(initial_dt)
        self.corrector_order = int(corrector_order)
        self.h_sub_ref = classic_h_sub_ref(self.q, self.m, self.G, self.initial_dt, self.split_n_max)
        self.top_dt = self.initial_dt                               # :148
        self.history = [self.s]
        self.force_evals = 0
        if (not skip_init_corrector and self.G != 0.0 and not self.adaptive_softening
                and not self.adaptive_timestep):                    # :150-157
            self.apply_corrector()

    # -- force -------------------------------------------------------------
    def eps_force(self):
        return math.sqrt(self.step_s2) if self.step_s2 > 0.0 else 0.0

    def accel(self):
        self.force_evals += 1
        if self.n < 2 or self.G == 0.0:
            return np.zeros_like(self.q)
        return accelerations(self.q, self.m, self.eps_force(), self.G)

    # -- corrector (ctor + snapshot half kick) -------------------------------
    def wh_interaction_accel(self):
        """whfast_scheme.py:39-69 (only used by the whfast corrector)."""
        n, m, pos, G, s2 = self.n, self.m, self.q, self.G, self.step_s2
        acc = np.zeros_like(pos)
        if n < 2:
            return acc
        cum = np.cumsum(m)
        jq, _ = self.to_jacobi()
        for i in range(2, n):
            rj = jq[i]
            rn2 = float(np.dot(rj, rj)) + s2
            if rn2 > 0:
                aj = G * cum[i - 1] * rj / (rn2 ** 1.5)
                for k in range(i):
                    acc[k] -= m[i] * aj * (m[k] / cum[i - 1])
                acc[i] += cum[i - 1] * aj
        for i in range(n):
            for j in range(i + 1, n):
                if not (i == 0 and j > 0):
                    dr = pos[j] - pos[i]
                    r2 = float(np.dot(dr, dr)) + s2
                    f = G * dr * r2 ** -1.5
                    acc[i] -= m[j] * f
                    acc[j] += m[i] * f
        return acc

    def apply_corrector(self):
        """integration_scheme_base.py:154-192 / whfast_scheme.py:95-123."""
        if self.corrector_order <= 0 or self.G == 0.0:
            return
        h_ref = abs(float(self.top_dt)) if self.top_dt else 0.0
        if not (math.isfinite(h_ref) and h_ref > 0.0):
            h_ref = abs(self.h_sub_ref)
        if not (math.isfinite(h_ref) and h_ref > 0.0):
            return
        if self.mode == "whfast":
            if self.n < 2:
                return
            self.v = self.v + 0.5 * h_ref * self.wh_interaction_accel()
            return
        if self.n < 1:
            return
        if self.n >= 2:
            self.v = self.v + (0.5 * h_ref) * self.accel()

    commit_state = apply_corrector     # simulation.py:319-322

    # -- Jacobi ----------------------------------------------------------------
    def to_jacobi(self):
        """simulation.py:487-508."""
        m, pos, vel, n = self.m, self.q, self.v, self.n
        jp = np.empty_like(pos)
        jv = np.empty_like(vel)
        R = m[0] * pos[0]
        V = m[0] * vel[0]
        M = m[0]
        jp[0] = pos[0]
        jv[0] = vel[0]
        for i in range(1, n):
            jp[i] = pos[i] - R / M
            jv[i] = vel[i] - V / M
            R = R + m[i] * pos[i]
            V = V + m[i] * vel[i]
            M = M + m[i]
        return jp, jv

    def from_jacobi(self, jp, jv):
        """simulation.py:509-534."""
        m, n = self.m, self.n
        pos = np.empty_like(jp)
        vel = np.empty_like(jv)
        pos[0] = jp[0]
        vel[0] = jv[0]
        R = m[0] * pos[0]
        V = m[0] * vel[0]
        M = m[0]
        for i in range(1, n):
            pos[i] = jp[i] + R / M
            vel[i] = jv[i] + V / M
            R = R + m[i] * pos[i]
            V = V + m[i] * vel[i]
            M = M + m[i]
        return pos, vel

    # -- integrators ---------------------------------------------------------------
    def _verlet_kernel(self, h):
        """integration_scheme_base.py:129-149 (two force evaluations, no FSAL)."""
        h2 = 0.5 * h
        a0 = self.accel()
        self.v = self.v + h2 * a0
        self.q = self.q + h * self.v
        a1 = self.accel()
        self.v = self.v + h2 * a1

    def _kepler_drift(self, dt):
        """whfast_scheme.py:22-37."""
        m = self.m
        cum = np.cumsum(m)
        jp, jv = self.to_jacobi()
        jp[0] = jp[0] + jv[0] * dt
        for i in range(1, self.n):
            mu = self.G * (cum[i - 1] + m[i])
            rn, vn = kepler_propagate(jp[i], jv[i], mu, dt)
            jp[i] = rn
            jv[i] = vn
        self.q, self.v = self.from_jacobi(jp, jv)

    def _wisdom_holman(self, h):
        """whfast_scheme.py:71-93 -- Kepler(h/2), FULL-force kick(h), Kepler(h/2)."""
        self._kepler_drift(0.5 * h)
        if self.n >= 2 and self.G != 0.0:
            self.v = self.v + h * self.accel()
        self._kepler_drift(0.5 * h)

    def n_sub_for(self, dt):
        """integrator.py:86-92."""
        h_sub = self.h_sub_ref
        return int(max(1, min(self.split_n_max, math.ceil(abs(dt) / h_sub))))

    def step(self, dt):
        """simulation.py:667-676 -> integrator.py:78-104."""
        if dt == 0.0 or self.n == 0:
            return
        dt = float(dt)
        self.top_dt = abs(dt)
        n_sub = self.n_sub_for(dt)
        h = dt / n_sub
        self.step_s2 = self.s ** 2                                   # begin_step (softening_manager.py:186-199)
        self.pending_energy_delta = 0.0
        self.history.append(self.s)
        if len(self.history) > 1024:
            self.history = self.history[-1024:]
        for _ in range(n_sub):
            if self.mode == "yoshida4":
                self._verlet_kernel(YOSHIDA_W1 * h)
                self._verlet_kernel(YOSHIDA_W2 * h)
                self._verlet_kernel(YOSHIDA_W1 * h)
            elif self.mode == "whfast":
                self._wisdom_holman(h)
            else:
                self._verlet_kernel(h)
            if self.adaptive_softening:                              # integrator.py:126-136, 204-225
                self.refresh_softening(self.softening_from_min_sep(self.min_separation()))
            if self.pending_energy_delta != 0.0:                     # commit_substep (softening_manager.py:246-257)
                self.softening_energy_delta = self.softening_energy_delta + self.pending_energy_delta
                self.pending_energy_delta = 0.0

    # -- classic adaptive softening (softening_manager.py:298-336, 423-471, 541-547) ------------------
    def min_separation(self):
        """simulation.py:659-665."""
        if self.n < 2:
            return float("inf")
        d = self.q[:, None, :] - self.q[None, :, :]
        d2 = (d ** 2).sum(axis=-1)
        np.fill_diagonal(d2, np.inf)
        return max(float(d2.min()) ** 0.5, 1e-12)

    def softening_from_min_sep(self, min_sep):
        if not math.isfinite(min_sep) or min_sep <= 0.0:
            return self.s
        proposed = max(self.min_softening, min_sep / self.softening_scale)
        proposed = min(proposed, 10.0 * self.s0)
        return max(self.s / 2.0, min(self.s * 2.0, proposed))       # _limited_softening, factor 2

    def energy_correction(self, eps_old, eps_new):
        """softening_manager.py:423-471 for the classic integrators (k_soft = 0: no spring term)."""
        if eps_old == eps_new:
            return 0.0
        dE = 0.0
        if self.n >= 2 and self.G != 0.0:
            diff = self.q[:, None, :] - self.q[None, :, :]
            r2 = np.einsum("ijk,ijk->ij", diff, diff, optimize=True)
            np.fill_diagonal(r2, np.inf)
            inv_old = 1.0 / np.sqrt(r2 + eps_old ** 2)
            inv_new = 1.0 / np.sqrt(r2 + eps_new ** 2)
            np.fill_diagonal(inv_old, 0.0)
            np.fill_diagonal(inv_new, 0.0)
            iu, ju = np.triu_indices(self.n, 1)
            dE += self.G * float(np.sum(self.m[iu] * self.m[ju] * (inv_new[iu, ju] - inv_old[iu, ju])))
        dE += (_barrier_energy(eps_new, self.min_softening, 10.0 * self.s0, self.k_wall, self.barrier_exponent)
               - _barrier_energy(eps_old, self.min_softening, 10.0 * self.s0, self.k_wall, self.barrier_exponent))
        return float(dE)

    def refresh_softening(self, eps_new):
        eps_old, eps_new = float(self.s), float(eps_new)
        dE = 0.0
        if math.isfinite(eps_old) and math.isfinite(eps_new):
            val = self.energy_correction(eps_old, eps_new)
            if math.isfinite(val):
                dE = val
        self.pending_energy_delta = self.pending_energy_delta + dE
        self.s = eps_new
        self.step_s2 = eps_new * eps_new
        # softening_manager.py:331-335 appends to `self.history`, a property that returns a COPY of the deque: the
        # append is lost, so only begin_step records history entries

    # -- snapshot / restore ------------------------------------------------------------
    def snapshot_restore(self):
        """simulation.py:324-395 + 399-484: snapshot() kicks self, restore() rebuilds the
        sub-step schedule from the snapshotted positions with both corrector and COM
        re-centring skipped."""
        self.commit_state()
        c = OracleSim(self.m.copy(), self.q.copy(), self.v.copy(), G=self.G, softening=self.history[0],
                      min_softening=(0.1 * self.history[0] if self.history[0] > 0 else 0.0),
                      integrator_mode=self.mode, skip_init_corrector=True, skip_cm_recenter=True,
                      initial_dt=self.initial_dt, split_n_max=self.split_n_max,
                      corrector_order=self.corrector_order, adaptive_softening=self.adaptive_softening,
                      adaptive_timestep=self.adaptive_timestep, softening_scale=self.softening_scale,
                      k_wall=self.k_wall, barrier_exponent=self.barrier_exponent)
        # simulation.py:473-482: after the state is restored the manager is reset to the ORIGINAL sim's `_epsilon`
        # (update_continuous), which for classic modes is the constructor softening -- an adaptive copy therefore
        # restarts its softening from s0 while keeping the restored history
        c.s = self.eps_attr
        c.step_s2 = c.s * c.s
        c.eps_attr = self.eps_attr
        c.history = list(self.history)
        c.top_dt = self.top_dt
        c.softening_energy_delta = self.softening_energy_delta
        return c


# --------------------------------------------------------------------------
# diagnostics  diagnostics.py
# --------------------------------------------------------------------------


def _kahan_hp(arr):
    """diagnostics.py:423-441."""
    a = np.asarray(arr, dtype=HP).ravel()
    s = HP(0.0)
    c = HP(0.0)
    for x in a:
        y = x - c
        t = s + y
        c = (t - s) - y
        s = t
    return s


def extended_hamiltonian_classic(m, q, v, eps, G):
    """diagnostics.py:457-549 for classic modes (k_soft = 0, pi = 0, no barrier):
    long-double Kahan T and V, each cast to fp64, then added."""
    m_hp = np.asarray(m, dtype=HP)
    v_hp = np.asarray(v, dtype=HP)
    v2 = np.sum(v_hp * v_hp, axis=1, dtype=HP)
    T_hp = _kahan_hp(HP(0.5) * m_hp * v2)
    n = len(m)
    e = HP(eps)
    if n >= 2 and G != 0.0:
        p = np.asarray(q, dtype=HP)
        diff = p[:, None, :] - p[None, :, :]
        r2 = np.sum(diff * diff, axis=-1, dtype=HP)
        iu = np.triu_indices(n, 1)
        r2iu = r2[iu] + e * e
        r2iu = np.where(r2iu > HP(0.0), r2iu, HP(1e-300))
        inv_r = HP(1.0) / np.sqrt(r2iu, dtype=HP)
        V_hp = HP(-G) * _kahan_hp(m_hp[iu[0]] * m_hp[iu[1]] * inv_r)
    else:
        V_hp = HP(0.0)
    return float(T_hp) + float(V_hp)


def angular_momentum(m, q, v):
    """diagnostics.py:553-557."""
    s = 0.0
    for i in range(len(m)):
        s += float(m[i]) * (float(q[i, 0]) * float(v[i, 1]) - float(q[i, 1]) * float(v[i, 0]))
    return s


def step_metrics_classic(m, q, v, eps, L0_first):
    """diagnostics.py:241-285 for classic modes (pi = 0, mu_soft = 1)."""
    com_drift = float(np.linalg.norm(np.sum(m[:, None] * q, axis=0)))
    J_eps = float(eps * 0.0 / 1.0)
    theta_e = math.atan2(0.0, 1.0 * eps) if (1.0 * eps) else float("nan")
    L_i = m * (q[:, 0] * v[:, 1] - q[:, 1] * v[:, 0])
    L_tot = float(np.sum(L_i))
    var_L = float(np.var(L_i))
    if L0_first is None:
        L0_first = L_tot
    if L0_first and L_tot:
        cos_theta = (L_tot * L0_first) / (abs(L_tot) * abs(L0_first))
    else:
        cos_theta = float("nan")
    return dict(com_drift=com_drift, J_eps=J_eps, theta_eps=theta_e, var_L=var_L,
                cos_theta=cos_theta, tr_hessian=0.0, L_tot=L_tot), L0_first


def megno_init_vectors(m, raw_r, raw_v):
    """evolution_features.py:37-44 -- mass-weighted mean removed, unit Frobenius norm."""
    m = np.asarray(m, dtype=float)
    dr = np.array(raw_r, dtype=float)
    dr -= np.sum(m[:, None] * dr, axis=0) / np.sum(m)
    dr /= np.linalg.norm(dr)
    dv = np.array(raw_v, dtype=float)
    dv -= np.sum(m[:, None] * dv, axis=0) / np.sum(m)
    dv /= np.linalg.norm(dv)
    return dr, dv


def compute_megno(sim, n_steps, dt, raw_r=None, raw_v=None):
    """evolution_features.py:34-66.  raw_r/raw_v are the two randn(n,2) draws; when None they
    are drawn from the global NumPy RNG in the reference's order (delta_r first)."""
    n = sim.n
    if raw_r is None:
        raw_r = np.random.randn(n, 2)
        raw_v = np.random.randn(n, 2)
    dr, dv = megno_init_vectors(sim.m, raw_r, raw_v)
    t = 0.0
    accum = 0.0
    for _ in range(n_steps):
        sim.step(dt)
        dr = dr + dv * dt
        da = variational_accel(sim.q, sim.m, sim.step_s2, dr, sim.G)
        dv = dv + da * dt
        t += dt
        nr = float(np.linalg.norm(dr))
        if nr < 1e-12:
            dr = dr / nr
            dv = dv / nr
            nr = 1.0
        nv = float(np.linalg.norm(dv))
        accum += (nv / nr) * t * dt
    Y = 2.0 * accum / t
    lyap = math.inf if Y == 0.0 else t / abs(Y)
    return float(Y), float(lyap), dr, dv


def _drift(a0, a1):
    """stability_analyzer.py:147-170."""
    if np.isfinite(a0) and abs(a0) > 0.0 and np.isfinite(a1):
        return abs((a1 - a0) / a0)
    if np.isfinite(a0) and np.isfinite(a1):
        return abs(a1 - a0)
    return float("inf")


def dynamical_features(sim) -> Dict[str, float]:
    """dynamical_features.py:27-155 (25 static features of the current state)."""
    m, q, v, n = sim.m, sim.q, sim.v, sim.n
    out: Dict[str, float] = {}
    out["total_mass"] = float(np.sum(m))
    out["mass_variance"] = float(np.var(m))
    out["mass_ratio_max"] = float(np.max(m) / np.min(m)) if np.min(m) > 0 else 1.0
    M = 0.0
    xs = ys = 0.0
    for i in range(n):
        M += float(m[i])
    for i in range(n):
        xs += float(m[i]) * float(q[i, 0])
        ys += float(m[i]) * float(q[i, 1])
    out["mass_center_offset"] = float(np.sqrt((xs / M) ** 2 + (ys / M) ** 2)) if M != 0.0 else 0.0
    dists = []
    rels = []
    for i in range(n):
        for j in range(i + 1, n):
            dx = q[j, 0] - q[i, 0]
            dy = q[j, 1] - q[i, 1]
            dists.append(np.sqrt(dx * dx + dy * dy))
            dvx = v[j, 0] - v[i, 0]
            dvy = v[j, 1] - v[i, 1]
            rels.append(np.sqrt(dvx * dvx + dvy * dvy))
    if dists:
        mn, mx = float(min(dists)), float(max(max(dists), 0.0))
        out["mean_separation"] = float(np.mean(dists))
        out["std_separation"] = float(np.std(dists))
    else:
        mn, mx = 0.0, 0.0
        out["mean_separation"] = 0.0
        out["std_separation"] = 0.0
    out["min_separation"] = mn
    out["max_separation"] = mx
    out["separation_ratio"] = mx / mn if mn > 0 else 1.0
    speeds = np.sqrt(np.sum(v ** 2, axis=1))
    out["mean_speed"] = float(np.mean(speeds))
    out["std_speed"] = float(np.std(speeds))
    out["max_speed"] = float(np.max(speeds))
    out["mean_relative_velocity"] = float(np.mean(rels)) if rels else 0.0
    out["max_relative_velocity"] = float(np.max(rels)) if rels else 0.0
    KE = 0.0
    for i in range(n):
        KE += 0.5 * float(m[i]) * (float(v[i, 0]) ** 2 + float(v[i, 1]) ** 2)
    PE = 0.0
    for i in range(n):
        for j in range(i + 1, n):
            dx = float(q[j, 0]) - float(q[i, 0])
            dy = float(q[j, 1]) - float(q[i, 1])
            PE -= sim.G * float(m[i]) * float(m[j]) / math.sqrt(dx * dx + dy * dy + sim.step_s2)
    E = KE + PE
    out["kinetic_energy"] = KE
    out["potential_energy"] = PE
    out["total_energy"] = E
    out["virial_ratio"] = 2 * KE / abs(PE) if PE != 0 else 0.0
    out["energy_per_mass"] = E / float(np.sum(m))
    out["is_bound"] = float(E < 0)
    L = angular_momentum(m, q, v)
    spec = [abs(float(m[i]) * (q[i, 0] * v[i, 1] - q[i, 1] * v[i, 0])) / m[i] for i in range(n)]
    out["total_angular_momentum"] = abs(L)
    out["mean_specific_angular_momentum"] = float(np.mean(spec))
    out["angular_momentum_variance"] = float(np.var(spec))
    hist = list(sim.history)
    out["softening_mean"] = float(np.mean(hist))
    out["softening_std"] = float(np.std(hist))
    return out


def run_stability_analysis(sim, n_steps=1000, dt=0.01, mode="core", raw_r=None, raw_v=None):
    """stability_analyzer.py:69-259 (+ batch_stability_analyzer.py:37-58 post-processing) for
    classic integrator modes.  `sim` is mutated by the snapshot kick exactly like the reference."""
    n_steps = max(1, int(n_steps))
    c = sim.snapshot_restore()
    eps = c.eps_attr              # diagnostics.py:474: eps = sim._epsilon, constant for classic modes
    E0 = extended_hamiltonian_classic(c.m, c.q, c.v, eps, c.G)
    if mode == "minimal":
        for _ in range(n_steps):
            c.step(dt)
        E1 = extended_hamiltonian_classic(c.m, c.q, c.v, c.eps_attr, c.G)
        d = _drift(E0, E1)
        return {"is_stable": float(d < 0.01), "energy_drift": d, "mode": "minimal"}
    L0 = angular_momentum(c.m, c.q, c.v)
    samples = {k: [] for k in ("com_drift", "J_eps", "theta_eps", "cos_theta", "var_L", "tr_hessian")}
    interval = max(1, n_steps // 100)
    Lfirst = None
    for i in range(n_steps):
        c.step(dt)
        if i % interval == 0:
            met, Lfirst = step_metrics_classic(c.m, c.q, c.v, c.eps_attr, Lfirst)
            for k in samples:
                samples[k].append(met[k])
    E1 = extended_hamiltonian_classic(c.m, c.q, c.v, c.eps_attr, c.G)
    L1 = angular_momentum(c.m, c.q, c.v)
    if mode == "full":
        n_samp = min(50, n_steps // 2)
        if n_samp > 0:
            megno, lyap, _, _ = compute_megno(c, min(100, n_samp), dt, raw_r, raw_v)
        else:
            megno, lyap = 2.0, float("inf")
    else:
        megno, lyap = 2.0, float("inf")
    ed = _drift(E0, E1)
    ld = _drift(L0, L1)
    com_mean = float(np.mean(samples["com_drift"]))
    res = {
        "is_stable": float((ed < 0.01) and (ld < 0.01) and (com_mean < 1.0) and (megno < 10.0)),
        "energy_drift": ed,
        "angular_momentum_drift": ld,
        "com_drift_mean": com_mean,
        "com_drift_max": float(np.max(samples["com_drift"])),
        "j_eps_mean": float(np.mean(samples["J_eps"])),
        "j_eps_std": float(np.std(samples["J_eps"])),
        "theta_eps_mean": float(np.mean(samples["theta_eps"])),
        "theta_eps_std": float(np.std(samples["theta_eps"])),
        "cos_theta_mean": float(np.mean(samples["cos_theta"])),
        "cos_theta_min": float(np.min(samples["cos_theta"])),
        "ang_mom_var_mean": float(np.mean(samples["var_L"])),
        "ang_mom_var_max": float(np.max(samples["var_L"])),
        "tidal_trace_mean": float(np.mean(samples["tr_hessian"])),
        "tidal_trace_max": float(np.max(samples["tr_hessian"])),
        "MEGNO": float(megno),
        "lyapunov_time": float(lyap),
        "mode": mode,
    }
    if mode == "full":
        for k, val in dynamical_features(sim).items():
            res["initial_" + k] = val
    if abs(res["energy_drift"]) > 10:
        res["is_stable"] = 0.0
        res["pathological_energy"] = True
    else:
        res["pathological_energy"] = False
    res["softening_policy"] = "static"
    return res
