"""Pair-kernel functions with the reference's signatures (forces.py, potential.py, tangent_map.py), evaluated by
`nb_pair_batched_f64` / `nb_variational_batched_f64`.  Single systems are a batch of one; pass 3-D arrays
`q[B,N,2]`, `m[B,N]` to evaluate a whole batch in one launch (extension).  N = 2..8 per system here; large N
goes through `LargeNSimulation`."""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import ensemble as E


def _prep(q, m):
    q = np.asarray(q, dtype=float)
    m = np.asarray(m, dtype=float)
    single = q.ndim == 2
    if single:
        q, m = q[None], m.reshape(1, -1)
    return q, m, single


def gravitational_force(q, m, eps: float = 0.0, G: float = 1.0):
    """forces.py:63-75: F_i = sum_j -G m_i m_j (r^2+eps^2)^-1.5 (q_i - q_j); zeros for N < 2 or G == 0."""
    qa, ma, single = _prep(q, m)
    if qa.shape[1] < 2 or G == 0.0:
        return np.zeros_like(np.asarray(q, dtype=float))
    acc, _, _ = E.pair_batched(qa, ma, eps, G, want_U=False, want_dV=False)
    F = acc.cpu().numpy() * ma[:, :, None]
    return F[0] if single else F


pairwise_force = gravitational_force


def softened_forces(q, m, G, eps):
    """forces.py:35-59 (same maths, zeros on malformed input)."""
    qa = np.asarray(q, dtype=float)
    ma = np.asarray(m, dtype=float)
    if qa.ndim != 2 or qa.shape[1] != 2 or ma.size != qa.shape[0] or qa.shape[0] < 2 or float(G) == 0.0:
        return np.zeros_like(qa, dtype=float)
    return gravitational_force(qa, ma, float(eps), float(G))


def dV_d_epsilon(q, m, eps, G: float = 1.0):
    """forces.py:77-112: G eps sum_{i<j} m_i m_j (r^2+eps^2)^-1.5 (0 if eps == 0)."""
    qa, ma, single = _prep(q, m)
    if qa.shape[1] < 2 or float(G) == 0.0:
        return 0.0 if single else np.zeros(qa.shape[0])
    _, _, dV = E.pair_batched(qa, ma, eps, G, want_acc=False, want_U=False)
    dV = dV.cpu().numpy()
    return float(dV[0]) if single else dV


def softened_potential(q, m, G, eps):
    """potential.py:23-64: -G sum_{i<j} m_i m_j / sqrt(r^2+eps^2)."""
    qa, ma, single = _prep(q, m)
    if qa.shape[1] < 2 or float(G) == 0.0:
        return 0.0 if single else np.zeros(qa.shape[0])
    _, U, _ = E.pair_batched(qa, ma, eps, G, want_acc=False, want_dV=False)
    U = U.cpu().numpy()
    return float(U[0]) if single else U


def dU_d_eps(q, m, G, eps):
    """potential.py:67-74."""
    return dV_d_epsilon(q, m, float(eps), float(G))


class TangentMap:
    """tangent_map.py:17-59."""

    def __init__(self, sim):
        self.sim = sim

    def variational_accel(self, delta_r):
        sim = self.sim
        delta_r = np.asarray(delta_r, dtype=float)
        if sim.n_bodies < 2 or sim.G == 0.0:
            return np.zeros_like(delta_r)
        da = E.variational_batched(sim._pos[None], sim._mass[None], float(sim.manager.step_s2), delta_r[None], sim.G,
                                   sim.device)
        return da.cpu().numpy()[0]
