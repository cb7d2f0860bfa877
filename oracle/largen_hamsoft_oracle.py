"""CPU oracle (test infrastructure, NOT a product path) for the large-N ham_soft variant.

The large-N path (nbodysimproject_b200/largen.py: LargeNHamSoftSimulation) runs the reference's ham_soft flow with
ONE documented change: grad eps* is always the analytic `_production_grad`
(hamsoft_eps_model.py:451-556), sign-aligned against softening.py:86-131 exactly as the reference's own fallback
branch does (hamsoft_eps_model.py:200-230), instead of the 4N-solve central difference, which is O(N^3).
This oracle is `HamSoftOracleSim` (bit-exact against the live reference, tests/golden/hamsoft.npz) with that branch
forced, and with the pair loops of `solve_hi` / `production_grad` restated as dense NumPy so that N ~ 10^2..10^3
finishes in seconds.  tests/test_oracle_golden.py checks the dense restatement against the loop version
(which is the pinned one) to 1e-12.  Parity status: pinned through that chain.
"""
from __future__ import annotations

import numpy as np

from .hamsoft_oracle import HamSoftOracleSim, legacy_grad


class LargeNHamSoftOracle(HamSoftOracleSim):
    def __init__(self, *a, dense=True, **kw):
        self.dense = dense
        super().__init__(*a, **kw)

    # -- dense restatement of hamsoft_eps_model.py:316-400
    def solve_hi(self, q):
        if not self.dense:
            return super().solve_hi(q)
        self.n_solves += 1
        m = self.m
        n = int(q.shape[0])
        eps_min, eps_max = self.eps_min, self.eps_max
        if eps_max < eps_min:
            eps_min, eps_max = eps_max, eps_min
        floor = max(eps_min, 1.0e-12)
        cap = max(floor, eps_max)
        h0 = float(self.eps)
        if not np.isfinite(h0) or h0 <= 0.0:
            h0 = 1.0
        h0 = min(max(h0, floor), cap)
        h = np.full(n, h0)
        d = q[:, None, :] - q[None, :, :]
        r = np.hypot(d[..., 0], d[..., 1])
        r2 = r * r
        off = ~np.eye(n, dtype=bool)
        it = 0
        self.sweeps = 0
        while it < 8:
            hj = np.maximum(h, 1.0e-12)
            c = 1.0 / (np.pi * hj * hj)
            W = c[:, None] * np.exp(-r2 / (hj * hj)[:, None])
            Sigma = np.sum(np.where(off, m[None, :] * W, 0.0), axis=1)
            Si = np.maximum(Sigma, 1.0e-30)
            hn = float(self.eta) * np.sqrt(m / Si)
            hn = np.where(np.isfinite(hn) & (hn > 0.0), hn, h)
            hn = np.clip(hn, floor, cap)
            changed = float(np.max(np.abs(hn - h) / np.maximum(h, 1.0e-12)))
            h = hn
            self.sweeps += 1
            if changed < 1.0e-6:
                break
            it += 1
        return h

    # -- dense restatement of hamsoft_eps_model.py:451-556
    def production_grad(self, q):
        if not self.dense:
            return super().production_grad(q)
        q = np.asarray(q, dtype=float)
        m, n = self.m, int(q.shape[0])
        a = self._alpha()
        h = self.solve_hi(q)
        floor = max(self.eps_min, 1.0e-12)
        hmin = max(1.0e-12, 0.1 * floor)
        t = -h / a
        tmax = float(np.max(t))
        ex = np.exp(t - tmax)
        den = float(np.sum(ex))
        if den <= 0.0 or not np.isfinite(den):
            return np.zeros((n, 2))
        w = ex / den
        hj = np.maximum(h, hmin)
        d = q[:, None, :] - q[None, :, :]
        r = np.hypot(d[..., 0], d[..., 1])
        r2 = r * r
        off = ~np.eye(n, dtype=bool)
        c = 1.0 / (np.pi * hj * hj)
        W = np.where(off, c[:, None] * np.exp(-r2 / (hj * hj)[:, None]), 0.0)
        Sigma = np.sum(m[None, :] * W, axis=1)
        Sd = np.sum(m[None, :] * (W * (-2.0 / hj[:, None] + 2.0 * r2 / (hj ** 3)[:, None])), axis=1)
        Si = np.maximum(Sigma, 1.0e-30)
        Om = 1.0 + hj * Sd / (2.0 * Si)
        Om = np.where(np.isfinite(Om) & (Om != 0.0), Om, 1.0)
        P = -hj / (2.0 * Si * Om)
        s_i = -w * P
        d2 = d[..., 0] ** 2 + d[..., 1] ** 2
        Wg = np.where(off, c[:, None] * np.exp(-d2 / (hj * hj)[:, None]), 0.0)
        coef = -2.0 * Wg / (hj * hj)[:, None]                       # [i, j], uses h_i
        K = s_i[:, None] * m[None, :] * coef                        # contribution of the (i, j) loop iteration
        g = np.einsum("ij,ijk->ik", K, d) - np.einsum("ij,ijk->jk", K, d)
        return np.where(np.isfinite(g), g, 0.0)

    # -- the reference's fallback branch, taken unconditionally (hamsoft_eps_model.py:200-230)
    def eps_star_and_grad(self, q, h_rel=1e-5, h_abs=1e-10):
        q = np.asarray(q, dtype=float)
        es = self.eps_target(q)
        g_use = self.production_grad(q)
        g_ref = legacy_grad(q, self.lam)
        if np.all(np.isfinite(g_ref)):
            dot = float(np.sum(g_use * g_ref))
            if np.isfinite(dot) and dot < 0.0:
                g_use = -g_use
        self.taps["fallback"] = True
        return float(es), g_use
