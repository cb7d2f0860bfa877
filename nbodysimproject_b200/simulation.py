"""NBodySimulation -- the reference's public facade (simulation.py:39-754) over the CUDA ensemble kernels.

The constructor arguments, attribute names (`_pos/_vel/_mass`, `pos/vel/mass/acc`, `manager`, `_integrator`,
`_epsilon/_pi`), the mode fall-backs (simulation.py:101-120), the constructor corrector half kick and the
snapshot/copy half kick (simulation.py:150-157, 319-326) are kept; every force evaluation, kick, drift and
Kepler solve runs on the GPU through the C ABI (`nb_ensemble_prepare_f64`, `nb_ensemble_run_f64`,
`nb_pair_batched_f64`).  Host NumPy arrays stay the user-visible source of truth (users mutate `_pos` in
place, like with the reference); they are uploaded per call.  `step_many(dt, n)` is an extension that keeps
the state in registers for n steps in ONE launch.
"""
from __future__ import annotations

import math
from collections import deque
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _lib as L
from . import ensemble as E

_ALLOWED_MODES = {"verlet", "yoshida4", "whfast", "ham_soft"}


@dataclass
class SimConfig:
    """sim_config.py:27-57."""
    safety_factor: float = 0.20
    theta_cap: float = 0.1
    theta_imp: float = 0.5
    k_soft: float = 1.0e3
    enable_runtime_guard: bool = False
    split_n_max: int = 50
    fast_float32: bool = False
    adaptive_timestep: bool = False
    adaptive_softening: bool = False
    softening_scale: float = 1.0
    integrator_mode: str = "ham_soft"
    use_energy_spring: bool = True
    use_soft_barrier: bool = True
    initial_dt: float = 0.01
    max_fraction_of_dt: float = 0.1
    corrector_order: int = 5
    disable_barrier: bool = False
    barrier_exponent: int = 5
    k_wall: float = 1.0e9
    n_wall: int = 4
    alpha: float | None = 0.1
    eta: float = 1.35
    guard_dt_ref: float = 1e-3
    energy_drift_abort_threshold: float = 1e-6
    ang_mom_drift_abort_threshold: float = 1e-5
    abort_on_violation: bool = True
    fixed_substeps: bool = True
    invariant_check_interval: int = 2000
    energy_tol_pref: float = 1e-8
    freeze_s_subsystem: bool = False

    def copy(self) -> "SimConfig":
        new = object.__new__(SimConfig)
        new.__dict__ = dict(self.__dict__)
        return new


@dataclass
class Body:
    """body.py -- plain value object accepted by NBodySimulation(bodies=[...])."""
    mass: float
    x: float
    y: float
    vx: float = 0.0
    vy: float = 0.0


class BodyView:
    """body_view.py:22-67 -- index proxy into the simulation arrays."""
    __slots__ = ("_sim", "_i")

    def __init__(self, sim, idx):
        self._sim, self._i = sim, int(idx)

    mass = property(lambda s: float(s._sim._mass[s._i]), lambda s, v: s._sim._mass.__setitem__(s._i, float(v)))
    x = property(lambda s: float(s._sim._pos[s._i, 0]), lambda s, v: s._sim._pos.__setitem__((s._i, 0), float(v)))
    y = property(lambda s: float(s._sim._pos[s._i, 1]), lambda s, v: s._sim._pos.__setitem__((s._i, 1), float(v)))
    vx = property(lambda s: float(s._sim._vel[s._i, 0]), lambda s, v: s._sim._vel.__setitem__((s._i, 0), float(v)))
    vy = property(lambda s: float(s._sim._vel[s._i, 1]), lambda s, v: s._sim._vel.__setitem__((s._i, 1), float(v)))

    def __repr__(self):
        return f"Body(mass={self.mass}, x={self.x}, y={self.y}, vx={self.vx}, vy={self.vy})"


class SofteningManager:
    """State holder part of softening_manager.py:37-372 (static softening and ham_soft paths; the classic
    adaptive refresh is SURVEY.md section 8f "next")."""

    def __init__(self, sim, softening: float, min_softening: float, history: int = 1024):
        self.sim = sim
        self.s0 = float(max(softening, min_softening))
        self._min_softening = float(min_softening)
        self.s = self.s0
        self.s2 = self.s * self.s
        self._step_s2 = self.s2
        self._history = deque([self.s], maxlen=int(history))
        self._pending_energy_delta = 0.0
        self._step_finished = True

    softening = property(lambda self: self.s)
    step_s2 = property(lambda self: self._step_s2)
    history = property(lambda self: list(self._history))

    def begin_step(self):                                    # softening_manager.py:186-199
        if self.sim._integrator_mode == "ham_soft":
            self.s = float(self.sim._epsilon)
        self._step_s2 = float(self.s) ** 2
        self._history.append(float(self.s))

    def finish_step(self):                                   # softening_manager.py:355-372
        if self.sim._integrator_mode == "ham_soft":
            self.s = float(self.sim._epsilon)
            self._step_s2 = self.s * self.s
        self._pending_energy_delta = 0.0

    def update_continuous(self, eps_new: float):             # softening_manager.py:338-353
        self.s = float(eps_new)
        self._step_s2 = self.s * self.s

    def debug_info(self):
        return dict(softening=self.s, step_s2=self._step_s2, history=list(self._history),
                    pending_energy_delta=0.0, last_dE=0.0, segments_used=0)


class _ClassicIntegrator:
    """Host-side state of integrator.py:28-104 for verlet / yoshida4 / whfast."""
    k_soft = 0.0
    mu_soft = 1.0
    chi_eps = 1.0

    def __init__(self, sim, split_n_max: int):
        self.sim = sim
        self.split_n_max = int(split_n_max)
        self._top_dt = None
        self._dt_prev = None
        self._eps_prev = None
        self._last_update_tick = 0
        self._cached_min_sep = None
        self._substeps_in_last_step = 0
        self._last_tr_hessian = 0.0
        self.h_sub_ref = float("nan")

    def n_sub_for(self, dt: float) -> int:
        return int(max(1, min(self.split_n_max, math.ceil(abs(dt) / self.h_sub_ref))))

    def compute_extended_hamiltonian(self) -> float:
        from .stability import Diagnostics
        return Diagnostics(self.sim, integrator=self).compute_extended_hamiltonian()


class NBodySimulation:
    def __init__(self, config=None, bodies=None, masses=None, positions=None, velocities=None, G: float = 1.0,
                 softening: float = 1e-3, min_softening: float = 0.0, adaptive: bool = False,
                 adaptive_timestep: bool = None, adaptive_softening: bool = None, skip_init_corrector: bool = False,
                 skip_cm_recenter: bool = False, integrator_mode: str | None = None, device=None):
        self.cfg = config.copy() if config else SimConfig()
        # reference test hooks that the GPU kernels do not implement (simulation.py:80-83, 159-162 float32 state arrays;
        # hamsoft_eps_model.py:82-89): reported, never silently ignored, never raised.  freeze_s_subsystem and
        # _validate_S_only (hamsoft_stepper.py:119-124, 270-284) ARE implemented: flags of the ham_soft parameter row
        for name in ("fast_float32", "use_legacy_eps_star", "fixed_eps_star"):
            if bool(getattr(self.cfg, name, False)):
                print(f"[nbodysimproject_b200] SimConfig.{name} is not supported by the fp64 GPU kernels: "
                      f"running the production path (NB_ERR_UNSUPPORTED at the C ABI)")
        self.device = device
        self.kepler_mode = "reference"       # or "exact" 
        if adaptive_timestep is not None:
            self._adaptive_timestep = bool(adaptive_timestep)
        elif adaptive is not None:
            self._adaptive_timestep = bool(adaptive)
        else:
            self._adaptive_timestep = bool(self.cfg.adaptive_timestep)
        self._adaptive_softening = bool(adaptive_softening) if adaptive_softening is not None else bool(self.cfg.adaptive_softening)
        if self._adaptive_softening and not self._adaptive_timestep:
            self._adaptive_timestep = True
        self.n_bodies = 0
        self._mass = np.empty(0)
        self._pos = np.empty((0, 2))
        self._vel = np.empty((0, 2))
        self._acc = np.empty((0, 2))
        if not self._build_state(bodies, masses, positions, velocities):
            self._disable_simulation()
            return
        min_softening = max(0.0, min_softening)
        if softening < 0.0:
            softening = min_softening
        if min_softening == 0.0 and softening > 0.0:
            min_softening = 0.1 * softening
        self._min_softening = float(min_softening)
        self._softening_scale = self.cfg.softening_scale
        if integrator_mode is not None:
            self.cfg.integrator_mode = str(integrator_mode)
        self._integrator_mode = self.cfg.integrator_mode
        self.G = float(G)
        if self.G == 0.0 and self._integrator_mode != "ham_soft":
            self._integrator_mode = "verlet"
        if self._integrator_mode == "whfast" and self.n_bodies > 0:
            if self._adaptive_softening:                                 # simulation.py:104-107
                self._integrator_mode = "verlet"
            elif np.max(self._mass) / np.sum(self._mass) < 0.2:
                self._integrator_mode = "verlet"
        self.manager = SofteningManager(self, softening, self._min_softening)
        self._max_softening = 10.0 * self.manager.s0
        self._epsilon = float(self.manager.s)
        self._pi = 0.0
        if self.manager.s > 0.0 and self._integrator_mode == "whfast":
            self._integrator_mode = "verlet"
        self.softening_energy_delta = 0.0
        self._has_integrated = False
        self._in_integration = False
        self._acc_cached = False
        self._status = 0
        # Body counts: 2..8 register-resident kernels, 9..64 one CTA (classic adaptive softening: one thread) per system,
        # every integrator mode.  The reference never raises (simulation.py:76-78): beyond 64 bodies the simulation is
        # reported and disabled (n_bodies = 0), like any other unusable input; LargeNSimulation is the fp32 large-N path.
        n_cap = 64
        if self.n_bodies > n_cap:
            print(f"[nbodysimproject_b200] {self.n_bodies} bodies in mode '{self._integrator_mode}': the fp64 kernels cover "
                  f"up to {n_cap}; simulation disabled -- use LargeNSimulation for large N")
            self._disable_simulation()
            return
        remove_com = not skip_cm_recenter
        if self._integrator_mode == "ham_soft":
            from .hamsoft import HamSoftIntegrator
            if remove_com:
                self._prepare(L.PREP_REMOVE_COM, 0.0)
            self._integrator = HamSoftIntegrator(self, split_n_max=self.cfg.split_n_max)
        else:
            self._integrator = _ClassicIntegrator(self, self.cfg.split_n_max)
            self._integrator._top_dt = getattr(self.cfg, "initial_dt", self.cfg.max_fraction_of_dt)
            flags = L.PREP_REMOVE_COM if remove_com else 0
            if (not skip_init_corrector and self.G != 0.0 and int(self.cfg.corrector_order) > 0
                    and not self._adaptive_softening and not self._adaptive_timestep):     # simulation.py:150-157
                flags |= L.PREP_CTOR_KICK
            self._prepare(flags, float(self._integrator._top_dt), schedule=True)

    # -- state --------------------------------------------------------------------------------
    def _build_state(self, bodies, masses, positions, velocities) -> bool:
        """simulation_state.py:98-144."""
        if bodies is None:
            if masses is None or positions is None:
                return False
            masses = list(masses)
            positions = list(positions)
            velocities = [] if velocities is None else list(velocities)
            if len(velocities) == 0:
                velocities = [(0.0, 0.0)] * len(masses)
            elif len(velocities) == 1 and len(masses) > 1:
                velocities = velocities * len(masses)
            if len(velocities) != len(masses):
                return False
            m = np.asarray(masses, dtype=np.float64)
            q = np.asarray(positions, dtype=np.float64).reshape(-1, 2)
            v = np.asarray(velocities, dtype=np.float64).reshape(-1, 2)
        else:
            m = np.array([b.mass for b in bodies], dtype=np.float64)
            q = np.array([(b.x, b.y) for b in bodies], dtype=np.float64).reshape(-1, 2)
            v = np.array([(b.vx, b.vy) for b in bodies], dtype=np.float64).reshape(-1, 2)
        if np.any(m <= 0) or not np.all(np.isfinite(m)):
            return False
        self.n_bodies = len(m)
        self._mass, self._pos, self._vel = m.copy(), q.copy(), v.copy()
        self._acc = np.zeros_like(self._pos)
        return True

    def _disable_simulation(self):
        self.n_bodies = 0
        self._mass = np.empty(0)
        self._pos = np.empty((0, 2))
        self._vel = np.empty((0, 2))
        self._acc = np.empty((0, 2))
        self.softening_energy_delta = 0.0
        self.G = getattr(self, "G", 1.0)
        self._integrator_mode = getattr(self, "_integrator_mode", self.cfg.integrator_mode)

    mass = property(lambda self: self._mass)
    pos = property(lambda self: self._pos)
    vel = property(lambda self: self._vel)
    acc = property(lambda self: self._acc)
    integrator_mode = property(lambda self: str(self._integrator_mode))
    soft = property(lambda self: self.manager.s)
    s = property(lambda self: self.manager.s)
    softening = property(lambda self: self.manager.softening)
    max_softening = property(lambda self: self._max_softening)
    adaptive_softening = property(lambda self: self._adaptive_softening)

    @pos.setter
    def pos(self, value):
        arr = np.asarray(value, dtype=np.float64).reshape(-1, 2)
        if arr.shape != self._pos.shape:
            print(f"shape mismatch when assigning to sim.pos: expected {self._pos.shape}, got {arr.shape}")
            return
        self._pos[...] = arr

    @vel.setter
    def vel(self, value):
        arr = np.asarray(value, dtype=np.float64).reshape(-1, 2)
        if arr.shape != self._vel.shape:
            print(f"shape mismatch when assigning to sim.vel: expected {self._vel.shape}, got {arr.shape}")
            return
        self._vel[...] = arr

    @property
    def bodies(self) -> List[BodyView]:
        return [BodyView(self, i) for i in range(self.n_bodies)]

    # -- GPU calls -------------------------------------------------------------------------------
    def _force_eps(self) -> float:
        if self._integrator_mode == "ham_soft":
            return float(self._epsilon)
        s2 = float(self.manager.step_s2)
        return math.sqrt(s2) if s2 > 0.0 else 0.0

    def _prepare(self, flags: int, kick_dt: float, schedule: bool = False):
        if self.n_bodies < 2:
            if schedule:
                self._integrator.h_sub_ref = float(self.cfg.initial_dt)
            return
        bk = E.DeviceBucket(self._mass[None], self._pos[None], self._vel[None], self._force_eps(), self.G,
                            "whfast" if self._integrator_mode == "whfast" else "verlet", self.device)
        bk.prepare(flags, kick_dt, float(self.cfg.initial_dt), float(self.cfg.initial_dt), int(self.cfg.split_n_max))
        if flags & (L.PREP_REMOVE_COM | L.PREP_CTOR_KICK | L.PREP_SNAPSHOT_KICK):
            self._vel[...] = bk.v.cpu().numpy()[0]
        if schedule:
            self._integrator.h_sub_ref = float(bk.h_sub_ref[0])

    def accelerations(self) -> np.ndarray:
        return self._accel()

    def _accel(self, *, pos=None, s2=None) -> np.ndarray:
        """simulation.py:539-581."""
        if self.n_bodies < 2 or self.G == 0.0:
            self._acc.fill(0.0)
            return self._acc
        if self._integrator_mode == "ham_soft" or s2 is None:
            eps = self._force_eps()
        else:
            eps = math.sqrt(s2) if s2 > 0.0 else 0.0
        p = self._pos if pos is None else np.asarray(pos, dtype=np.float64)
        acc, _, _ = E.pair_batched(p[None], self._mass[None], eps, self.G, self.device, want_U=False, want_dV=False)
        self._acc[:] = acc.cpu().numpy()[0]
        self._last_force_eps = float(eps)
        return self._acc

    _compute_accelerations = _accel

    def step(self, dt: float) -> None:
        """simulation.py:667-676."""
        self.step_many(dt, 1)

    def step_many(self, dt: float, n_steps: int) -> None:
        """n_steps x step(dt) in one launch (extension; identical results to calling step() n times)."""
        if dt == 0.0 or self.n_bodies == 0 or n_steps <= 0:
            return
        dt = float(dt)
        if self._integrator_mode == "ham_soft":
            self._integrator.step_many(dt, int(n_steps))
        else:
            integ = self._integrator
            integ._top_dt = abs(dt)
            if self._adaptive_softening and self.n_bodies >= 2 and self.G != 0.0:
                # classic adaptive softening: epsilon follows the minimum separation after every sub-step
                # (integrator.py:126-136, 204-225; softening_manager.py:298-336, 423-471, 541-547)
                mgr = self.manager
                q, v, hist, dE, st = E.advance_bucket_adaptive(
                    self._mass[None], self._pos[None], self._vel[None], mgr.s, mgr.s0, self._min_softening,
                    float(self._softening_scale), np.array([integ.h_sub_ref]), self.G, self._integrator_mode, dt,
                    int(n_steps), integ.split_n_max, float(getattr(self.cfg, "k_wall", 1.0e9)),
                    int(getattr(self.cfg, "barrier_exponent", 5)), self.softening_energy_delta, self.device)
                self._pos[...] = q[0]
                self._vel[...] = v[0]
                self._status |= int(st[0])
                # begin_step records the softening each macro step STARTS with (softening_manager.py:186-199)
                starts = [mgr.s] + [float(x) for x in hist[0][:-1]]
                for x in starts[-1024:]:
                    mgr._history.append(x)
                mgr.s = float(hist[0][-1])
                mgr._step_s2 = mgr.s * mgr.s
                self.softening_energy_delta = float(dE[0])
                integ._substeps_in_last_step = integ.n_sub_for(dt)
                mgr.finish_step()
                self._has_integrated = True
                self._acc_cached = False
                self._last_dt = dt
                return
            for _ in range(min(int(n_steps), 1024)):
                self.manager.begin_step()
            if self.n_bodies >= 2:
                q, v, st = E.advance_bucket(self._mass[None], self._pos[None], self._vel[None], self._force_eps(),
                                            np.array([integ.h_sub_ref]), self.G, self._integrator_mode, dt,
                                            int(n_steps), integ.split_n_max, self.device,
                                            kepler_exact=(self.kepler_mode == "exact"))
                self._pos[...] = q[0]
                self._vel[...] = v[0]
                self._status |= int(st[0])
            else:
                self._pos += dt * n_steps * self._vel     # a single free body drifts
            integ._substeps_in_last_step = integ.n_sub_for(dt)
            self.manager.finish_step()
        self._has_integrated = True
        self._acc_cached = False
        self._last_dt = dt

    # -- mode / snapshot -------------------------------------------------------------------------------
    def set_integrator_mode(self, mode: str) -> None:
        """simulation.py:281-300."""
        if self.G == 0.0:
            mode = "verlet"
        if mode not in _ALLOWED_MODES:
            return
        self._integrator_mode = mode
        if mode == "ham_soft":
            from .hamsoft import HamSoftIntegrator
            self._adaptive_softening = False
            self._integrator = HamSoftIntegrator(self, split_n_max=self.cfg.split_n_max)
        else:
            self._integrator = _ClassicIntegrator(self, self.cfg.split_n_max)
            self._prepare(0, 0.0, schedule=True)

    def get_integrator_name(self) -> str:
        return self._integrator_mode

    def get_current_softening_squared(self) -> float:
        return self.manager.step_s2

    def commit_state(self) -> None:
        """simulation.py:319-322 -> apply_corrector: the snapshot/copy half kick (no-op for ham_soft,
        hamiltonian_softening_integrator.py:753-754)."""
        if self.n_bodies == 0 or self._integrator_mode == "ham_soft":
            return
        if self.G == 0.0 or int(self.cfg.corrector_order) <= 0:
            return
        top = self._integrator._top_dt
        h_ref = abs(float(top)) if isinstance(top, (int, float, np.floating)) and np.isfinite(top) and top != 0 else abs(self._integrator.h_sub_ref)
        if not (np.isfinite(h_ref) and h_ref > 0.0):
            return
        self._prepare(L.PREP_SNAPSHOT_KICK, h_ref)

    def snapshot(self) -> dict:
        """simulation.py:324-395 (same keys; mutates velocities through commit_state like the reference)."""
        self.commit_state()
        mgr = self.manager
        soft_state = {"s0": mgr.s0, "s": mgr.s, "s2": mgr.s2, "step_s2": mgr._step_s2, "_step_s2": mgr._step_s2,
                      "min_softening": self._min_softening, "_pending_energy_delta": mgr._pending_energy_delta,
                      "_history": list(mgr._history), "_step_finished": mgr._step_finished}
        print("[snapshot] softening_mgr_state keys included: " + ", ".join(sorted(soft_state.keys())))
        integ = self._integrator
        int_state = {"dt_prev": integ._dt_prev, "eps_prev": integ._eps_prev, "_top_dt": integ._top_dt,
                     "_last_update_tick": integ._last_update_tick, "_cached_min_sep": integ._cached_min_sep,
                     "k_soft": getattr(integ, "k_soft", 0.0), "mu_soft": getattr(integ, "mu_soft", 1.0)}
        flags = {"_acc_cached": self._acc_cached, "_in_integration": self._in_integration,
                 "softening_energy_delta": self.softening_energy_delta,
                 "_adaptive_timestep": self._adaptive_timestep, "_adaptive_softening": self._adaptive_softening,
                 "_epsilon": self._epsilon, "_pi": self._pi}
        return {
            "masses": self._mass, "positions": self._pos, "velocities": self._vel, "softening": soft_state["s"],
            "softening_s2": soft_state["s2"], "pending_energy": self.softening_energy_delta,
            "integrator_state": int_state, "softening_mgr_state": soft_state, "sim_state": flags,
            "cfg": self.cfg.copy(), "has_integrated": bool(self._has_integrated),
            "sim": {"masses": self._mass, "positions": self._pos, "velocities": self._vel, "flags": flags},
            "integrator": int_state, "softening_mgr": soft_state, "acc": self._acc,
        }

    @classmethod
    def restore(cls, state):
        """simulation.py:399-484 + simulation_state.py:231-280."""
        cfg_in = state.get("cfg", state.get("sim", {}).get("cfg"))
        cfg = cfg_in.copy() if cfg_in else SimConfig()
        sim_data = state.get("sim", state)
        soft = state.get("softening_mgr_state", state.get("softening_mgr", {}))
        flags = state.get("sim_state", sim_data.get("flags", {}))
        hist = soft.get("_history")
        s0 = None
        if isinstance(hist, (list, tuple)) and len(hist) > 0 and isinstance(hist[0], (int, float, np.floating)) \
                and np.isfinite(hist[0]):
            s0 = float(hist[0])
        if s0 is None:
            val = state.get("softening", soft.get("s", 1e-3))
            s0 = float(val) if isinstance(val, (int, float, np.floating)) else 1e-3
        ms = state.get("min_softening")
        min_snap = float(ms) if isinstance(ms, (int, float, np.floating)) else (0.1 * s0 if s0 > 0.0 else 0.0)
        sim = cls(config=cfg, masses=sim_data["masses"], positions=sim_data["positions"],
                  velocities=sim_data["velocities"], softening=s0, min_softening=min_snap,
                  adaptive_timestep=bool(flags.get("_adaptive_timestep", False)),
                  adaptive_softening=bool(flags.get("_adaptive_softening", False)), skip_init_corrector=True,
                  skip_cm_recenter=True, integrator_mode=getattr(cfg, "integrator_mode", None))
        sim._mass = np.array(sim_data["masses"], dtype=np.float64)
        sim._pos = np.array(sim_data["positions"], dtype=np.float64)
        sim._vel = np.array(sim_data["velocities"], dtype=np.float64)
        sim._acc = np.array(state["acc"], dtype=np.float64) if "acc" in state else np.zeros_like(sim._pos)
        mgr = sim.manager
        mgr.s = soft.get("s", mgr.s)
        mgr.s2 = soft.get("s2", mgr.s2)
        mgr._step_s2 = soft.get("_step_s2", mgr._step_s2)
        if hist is not None:
            mgr._history = deque(list(hist), maxlen=mgr._history.maxlen)
        sim._epsilon = float(flags.get("_epsilon", mgr.s))
        mgr.update_continuous(sim._epsilon)
        sim._pi = float(flags.get("_pi", 0.0))
        sim.softening_energy_delta = flags.get("softening_energy_delta", state.get("pending_energy", 0.0))
        sim._has_integrated = bool(state.get("has_integrated", False))
        int_state = state.get("integrator_state", state.get("integrator", {}))
        integ = sim._integrator
        if sim._integrator_mode == "ham_soft":
            integ.restore_params(int_state)
        else:
            integ.k_soft = float(int_state.get("k_soft", integ.k_soft))
            integ.mu_soft = float(int_state.get("mu_soft", integ.mu_soft))
        integ._top_dt = int_state.get("_top_dt")
        mgr.update_continuous(float(flags.get("_epsilon", mgr.s)))
        sim._max_softening = 10.0 * float(mgr.s0)
        return sim

    def copy(self, *, deep: bool = True):
        if not deep:
            return self
        self.commit_state()
        return NBodySimulation.restore(self.snapshot())

    __copy__ = lambda self: self.copy(deep=True)
    __deepcopy__ = lambda self, memo=None: self.copy(deep=True)

    def _get_min_separation(self) -> float:
        if self.n_bodies < 2:
            return float("inf")
        d = self._pos[:, None, :] - self._pos[None, :, :]
        d2 = (d ** 2).sum(axis=-1)
        np.fill_diagonal(d2, np.inf)
        return max(float(d2.min()) ** 0.5, 1e-12)

    def to_jacobi(self):
        """simulation.py:487-508 (host helper; the whfast kernel has its own register-resident version)."""
        m, pos, vel = self._mass, self._pos, self._vel
        jp, jv = np.empty_like(pos), np.empty_like(vel)
        R, V, M = m[0] * pos[0], m[0] * vel[0], m[0]
        jp[0], jv[0] = pos[0], vel[0]
        for i in range(1, len(m)):
            jp[i] = pos[i] - R / M
            jv[i] = vel[i] - V / M
            R = R + m[i] * pos[i]
            V = V + m[i] * vel[i]
            M = M + m[i]
        return jp, jv

    def set_softening_bounds(self, eps_min: float, eps_max: float, *, clamp_epsilon: bool = True,
                             reset_pi_on_clamp: bool = True) -> None:
        """simulation.py:679-728."""
        a, b = float(eps_min), float(eps_max)
        a = a if np.isfinite(a) else 0.0
        b = b if np.isfinite(b) else a
        if b < a:
            a, b = b, a
        a = max(a, 0.0)
        self._min_softening, self._max_softening = a, b
        if clamp_epsilon:
            e = float(self._epsilon)
            new = min(max(e, a), b)
            if new != e:
                self._epsilon = new
                if reset_pi_on_clamp:
                    self._pi = -float(self._pi)
                self.manager.update_continuous(new)
