"""ham_soft host mirror: parameter table, constructor calibration and stepping through the CUDA kernels.

Mirrors the reference's HamiltonianSofteningIntegrator surface that other code touches
(hamiltonian_softening_integrator.py: k_soft, mu_soft, k_wall, barrier_policy, _frozen_n_sub, _omega_spr0,
_eps_model._alpha_run, eps_star_and_grad, compute_extended_hamiltonian, step) while every number is produced
on the GPU: `nb_hamsoft_setup_f64` (calibration + frozen schedule), `nb_ensemble_run_f64(mode=ham_soft)`
(Strang steps), `nb_hamsoft_probe_f64` (eps*, grad eps*, H_ext).
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import ensemble as E

P = {name: i for i, name in enumerate(L.HS_PARAMS)}


def default_params(cfg, softening, min_softening, B=1):
    """One hs_params row per system from SimConfig defaults (sim_config.py:27-57, simulation.py:88-114)."""
    softening = np.broadcast_to(np.asarray(softening, dtype=np.float64), (B,))
    min_softening = np.broadcast_to(np.asarray(min_softening, dtype=np.float64), (B,))
    hs = np.zeros((B, L.N_HS))
    s0 = np.maximum(softening, min_softening)
    disabled = bool(getattr(cfg, "disable_barrier", False))
    soft = bool(getattr(cfg, "use_soft_barrier", True)) and not disabled
    alpha = getattr(cfg, "alpha", 0.1)
    hs[:, P["k_soft"]] = float(getattr(cfg, "k_soft", 1.0e3))
    hs[:, P["mu_soft"]] = 1.0
    hs[:, P["eps_min"]] = min_softening
    hs[:, P["eps_max"]] = 10.0 * s0
    hs[:, P["alpha_run"]] = float(alpha) if isinstance(alpha, (int, float)) and alpha > 0 else 0.0
    hs[:, P["k_wall"]] = float(getattr(cfg, "k_wall", 1.0e9))
    hs[:, P["barrier_n"]] = int(getattr(cfg, "barrier_exponent", 5))
    hs[:, P["eta"]] = float(getattr(cfg, "eta", 1.35))
    hs[:, P["j_max_cap"]] = float(getattr(cfg, "j_max_cap", 0.02))
    hs[:, P["lambda"]] = float(getattr(cfg, "lambda_softening", 0.3))
    # hamiltonian_softening_integrator.py:96-108: 0 soft barrier, 1 reflection (fold), 2 barrier disabled
    hs[:, P["policy"]] = 0.0 if soft else (2.0 if disabled else 1.0)
    hs[:, P["theta_imp"]] = float(getattr(cfg, "theta_imp", 0.5))
    hs[:, P["theta_cap"]] = float(getattr(cfg, "theta_cap", 0.1))
    hs[:, P["chi_pi"]] = float(getattr(cfg, "chi_pi", 0.2))
    hs[:, P["omega_spr0"]] = 0.0
    hs[:, P["s0"]] = s0
    # reference test hooks (hamsoft_stepper.py:119-124, 270-284, 592-600)
    hs[:, P["flags"]] = ((L.HS_FLAG_FREEZE_S if bool(getattr(cfg, "freeze_s_subsystem", False)) else 0)
                         | (L.HS_FLAG_S_ONLY if bool(getattr(cfg, "_validate_S_only", False)) else 0))
    return hs, s0


class HamSoftBucket:
    """B same-N systems in ham_soft mode, device resident."""

    def __init__(self, m, q, v, hs_params, eps_pi, G=1.0, device=None):
        torch = L.require_cuda()
        self.torch = torch
        self.bk = E.DeviceBucket(m, q, v, 0.0, G, "ham_soft", device)
        dev = self.bk.device
        self.hs = torch.as_tensor(np.ascontiguousarray(hs_params, dtype=np.float64)).to(dev)
        self.eps_pi = torch.as_tensor(np.ascontiguousarray(eps_pi, dtype=np.float64)).to(dev)
        self.n_sub = torch.ones((self.bk.B,), dtype=torch.int32, device=dev)
        self.bk.n_sub = self.n_sub

    def setup(self, calibrate: bool, freeze_dt=None):
        bk, torch = self.bk, self.torch
        flags = (1 if calibrate else 0) | (2 if freeze_dt is not None else 0)
        with torch.cuda.device(bk.device):
            L.check(L.load().nb_hamsoft_setup_f64(L.ptr(bk.m), L.ptr(bk.q), bk.G, bk.B, bk.N, flags,
                                                  float(freeze_dt or 0.0), L.ptr(self.hs), L.ptr(self.eps_pi),
                                                  L.ptr(self.n_sub), L.stream_ptr()), "nb_hamsoft_setup_f64")

    def bump_mu(self, dt):
        """hamiltonian_softening_integrator.py:232-242 (the kernel applies the same floor internally)."""
        k = self.hs[:, P["k_soft"]]
        mu_macro = k * (abs(float(dt)) / self.hs[:, P["theta_imp"]]) ** 2
        mu = self.hs[:, P["mu_soft"]]
        self.hs[:, P["mu_soft"]] = self.torch.where((k > 0) & (mu < mu_macro), mu_macro, mu)

    def sort(self):
        """n_sub-sorted launch order: systems that share a warp (N <= 4) then have equal sub-step counts almost always
        (a warp runs to the larger count of its two systems).  Results do not depend on it."""
        self.bk.n_sub = self.n_sub
        self.bk.sort()

    def run(self, dt, n_steps, sample_interval=0, n_megno=0, raw_dr=None, raw_dv=None, flags=L.RUN_WRITE_STATE,
            want_dyn=False, work=None, t_main=None):
        dyn = self.bk.run(dt, n_steps, sample_interval, n_megno, raw_dr, raw_dv, flags, want_dyn,
                          eps_pi=self.eps_pi, hs_params=self.hs, work=work, t_main=t_main)
        if n_steps + n_megno > 0:
            self.bump_mu(dt)
        return dyn

    def probe(self):
        bk, torch = self.bk, self.torch
        out = torch.empty((bk.B, 3 + 2 * bk.N), dtype=torch.float64, device=bk.device)
        with torch.cuda.device(bk.device):
            L.check(L.load().nb_hamsoft_probe_f64(L.ptr(bk.m), L.ptr(bk.q), L.ptr(bk.v), bk.G, bk.B, bk.N,
                                                  L.ptr(self.eps_pi), L.ptr(self.hs), L.ptr(out), L.stream_ptr()),
                    "nb_hamsoft_probe_f64")
        o = out.cpu().numpy()
        return o[:, 0], o[:, 1], o[:, 2] != 0, o[:, 3:].reshape(bk.B, bk.N, 2)


class _EpsModelView:
    def __init__(self, owner):
        self._owner = owner

    @property
    def _alpha_run(self):
        return float(self._owner._hs[0, P["alpha_run"]])

    def eps_target(self, q=None):
        return self._owner.eps_star_and_grad(q)[0]


class HamSoftIntegrator:
    """Host-side state of one ham_soft simulation (the facade's `sim._integrator`)."""

    def __init__(self, sim, split_n_max: int = 50):
        self.sim = sim
        self.split_n_max = int(split_n_max)
        self._top_dt = None
        self._dt_prev = None
        self._eps_prev = None
        self._last_update_tick = 0
        self._cached_min_sep = None
        self._substeps_in_last_step = 0
        self._total_substeps_in_last_step = 0
        self._last_tr_hessian = 0.0
        self.chi_eps = 1.0
        self.h_sub_ref = float("nan")
        hs, _ = default_params(sim.cfg, sim.manager.s0, sim._min_softening)
        self._hs = hs
        self._eps_model = _EpsModelView(self)
        self.barrier_policy = "soft" if hs[0, P["policy"]] == 0.0 else "reflection"
        dt0 = float(getattr(sim.cfg, "initial_dt", 1.0e-2))
        if sim.n_bodies >= 2:
            b = self._bucket()
            b.setup(calibrate=True, freeze_dt=dt0)
            self._pull(b)
            self._frozen_n_sub = int(b.n_sub[0])
        else:
            self._frozen_n_sub = 1
        self._macro_dt_frozen = abs(dt0)
        sim._min_softening = float(self._hs[0, P["eps_min"]])
        sim.manager.update_continuous(sim._epsilon)
        self.h_sub_ref = abs(dt0) / self._frozen_n_sub

    # -- parameter views ------------------------------------------------------------------------
    k_soft = property(lambda s: float(s._hs[0, P["k_soft"]]), lambda s, v: s._hs.__setitem__((0, P["k_soft"]), float(v)))
    mu_soft = property(lambda s: float(s._hs[0, P["mu_soft"]]),
                       lambda s, v: s._hs.__setitem__((0, P["mu_soft"]), float(v) if float(v) != 0.0 else 1.0))
    k_wall = property(lambda s: float(s._hs[0, P["k_wall"]]), lambda s, v: s._hs.__setitem__((0, P["k_wall"]), float(v)))
    _omega_spr0 = property(lambda s: float(s._hs[0, P["omega_spr0"]]))
    epsilon = property(lambda s: float(s.sim._epsilon))
    pi = property(lambda s: float(s.sim._pi))

    def _barrier_n(self):
        return int(self._hs[0, P["barrier_n"]])

    def restore_params(self, int_state):
        self.k_soft = float(int_state.get("k_soft", self.k_soft))
        self.mu_soft = float(int_state.get("mu_soft", self.mu_soft))

    # -- device round trips ------------------------------------------------------------------------------
    def _bucket(self):
        sim = self.sim
        self._hs[0, P["eps_min"]] = float(sim._min_softening)
        self._hs[0, P["eps_max"]] = float(sim._max_softening)
        return HamSoftBucket(sim._mass[None], sim._pos[None], sim._vel[None], self._hs,
                             np.array([[sim._epsilon, sim._pi]]), sim.G, sim.device)

    def _pull(self, b):
        self._hs = b.hs.cpu().numpy()
        ep = b.eps_pi.cpu().numpy()
        self.sim._epsilon, self.sim._pi = float(ep[0, 0]), float(ep[0, 1])

    def strang_substeps(self, dt: float) -> int:
        """hamiltonian_softening_integrator.py:781-888: reuse the frozen n_sub within 1 % of the frozen dt."""
        dt_abs = abs(float(dt))
        prev = self._macro_dt_frozen
        if not (prev > 0.0 and abs(dt_abs - prev) / prev <= 0.01) and self.sim.n_bodies >= 2:
            b = self._bucket()
            b.bump_mu(dt_abs)
            b.setup(calibrate=False, freeze_dt=dt_abs)
            self._frozen_n_sub = int(b.n_sub[0])
            self._macro_dt_frozen = dt_abs
        return int(self._frozen_n_sub)

    def step_many(self, dt: float, n_steps: int):
        sim = self.sim
        self._top_dt = abs(dt)
        n_pred = max(1, self.strang_substeps(dt))
        for _ in range(min(n_steps, 1024)):
            sim.manager.begin_step()
        if sim.n_bodies >= 2:
            b = self._bucket()
            b.n_sub[:] = n_pred
            b.run(dt, n_steps)
            self._pull(b)
            sim._pos[...] = b.bk.q.cpu().numpy()[0]
            sim._vel[...] = b.bk.v.cpu().numpy()[0]
            sim._status |= int(b.bk.status[0])
        else:
            sim._pos += dt * n_steps * sim._vel
        self._substeps_in_last_step = n_pred
        self._total_substeps_in_last_step = n_pred
        sim.manager.finish_step()

    def step(self, dt: float):
        self.step_many(dt, 1)

    def eps_star_and_grad(self, q=None):
        """hamiltonian_softening_integrator.py:588-627."""
        sim = self.sim
        b = self._bucket()
        if q is not None:
            b.bk.q.copy_(self_torch(b).as_tensor(np.asarray(q, dtype=np.float64)[None]).to(b.bk.device))
        es, _, _, g = b.probe()
        return float(es[0]), g[0]

    def _eps_target(self, q=None, **kw):
        return self.eps_star_and_grad(q)[0]

    def compute_extended_hamiltonian(self) -> float:
        _, H, _, _ = self._bucket().probe()
        return float(H[0])


def self_torch(b):
    return b.torch
