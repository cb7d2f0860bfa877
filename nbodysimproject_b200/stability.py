"""Diagnostics and stability analysis on the GPU ensemble kernels, with the reference's classes and columns.

  Diagnostics               diagnostics.py:27-583        (energies, momenta, step_metrics inputs)
  DynamicalFeatures         dynamical_features.py:18-155 (25 static features -> prepare kernel)
  EvolutionFeatures         evolution_features.py:24-87  (MEGNO -> megno phase of the run kernel)
  StabilityAnalyzer         stability_analyzer.py:31-259 (one system = a batch of one)
  BatchStabilityAnalyzer    batch_stability_analyzer.py:29-102 (the batched entry point: ONE launch sequence per
                            (N, integrator mode, G) bucket instead of a Python loop over systems)

Feature rows keep the reference's keys and order (46 columns in 'full' mode).  Tangent vectors for MEGNO are drawn
on the host from the global NumPy RNG in the reference's order (two randn(n,2) per system, in batch order).
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np

from . import _lib as L
from . import ensemble as E

_DYN_KEYS = L.DYN_COLUMNS[:17]


def _mode_of(sim) -> str:
    return sim._integrator_mode


class Diagnostics:
    def __init__(self, simulation, integrator=None):
        self.sim = simulation
        self._integ = integrator if integrator is not None else getattr(simulation, "_integrator", None)

    # cheap O(N) host reductions on the user-visible arrays (diagnostics.py:63-67, 553-583)
    def kinetic_energy(self):
        s = 0.0
        for b in self.sim.bodies:
            s += 0.5 * b.mass * (b.vx * b.vx + b.vy * b.vy)
        return s

    def angular_momentum(self):
        s = 0.0
        for b in self.sim.bodies:
            s += b.mass * (b.x * b.vy - b.y * b.vx)
        return s

    def linear_momentum(self):
        px = py = 0.0
        for b in self.sim.bodies:
            px += b.mass * b.vx
            py += b.mass * b.vy
        return px, py

    def center_of_mass(self):
        M = sum(b.mass for b in self.sim.bodies)
        if M == 0.0:
            return (0.0, 0.0), (0.0, 0.0)
        xs = sum(b.mass * b.x for b in self.sim.bodies)
        ys = sum(b.mass * b.y for b in self.sim.bodies)
        px, py = self.linear_momentum()
        return (xs / M, ys / M), (px / M, py / M)

    # pair sums go to the GPU
    def potential_energy(self):
        """diagnostics.py:69-78: -G sum m_i m_j / sqrt(r^2 + manager.step_s2)."""
        sim = self.sim
        if sim.n_bodies < 2:
            return 0.0
        s2 = float(sim.manager.step_s2)
        _, U, _ = E.pair_batched(sim._pos[None], sim._mass[None], math.sqrt(s2) if s2 > 0 else 0.0, sim.G, sim.device,
                                 want_acc=False, want_dV=False)
        return float(U[0])

    def compute_extended_hamiltonian(self) -> float:
        """diagnostics.py:457-549 (double-double T + V on the device; + spring / barrier terms for ham_soft)."""
        sim = self.sim
        if _mode_of(sim) == "ham_soft":
            return sim._integrator.compute_extended_hamiltonian()
        if sim.n_bodies < 2:
            return self.kinetic_energy()
        bk = E.DeviceBucket(sim._mass[None], sim._pos[None], sim._vel[None], float(sim._epsilon), sim.G, "verlet", sim.device)
        dyn = bk.run(0.01, 0, 0, 0, flags=L.RUN_ENERGY).cpu().numpy()[0]
        return float(dyn[L.DYN_COLUMNS.index("_E0")])

    energy = compute_extended_hamiltonian

    def step_metrics(self, megno_slope_history=None) -> dict:
        """diagnostics.py:241-285 for the current state (host reductions; the batched path samples in-kernel)."""
        sim, m, pos, vel = self.sim, self.sim._mass, self.sim._pos, self.sim._vel
        eps, pi = sim._epsilon, sim._pi
        mu = getattr(sim._integrator, "mu_soft", 1.0)
        L_i = m * (pos[:, 0] * vel[:, 1] - pos[:, 1] * vel[:, 0])
        L_tot = float(np.sum(L_i))
        if not hasattr(self, "_L0"):
            self._L0 = L_tot
        cos_theta = (L_tot * self._L0) / (abs(L_tot) * abs(self._L0)) if (self._L0 and L_tot) else float("nan")
        return dict(com_drift=float(np.linalg.norm(np.sum(m[:, None] * pos, axis=0))), J_eps=float(eps * pi / mu),
                    L_tot=L_tot, var_L=float(np.var(L_i)), cos_theta=cos_theta, tr_hessian=0.0,
                    megno_slope_med=float(np.median(megno_slope_history)) if megno_slope_history else float("nan"),
                    theta_eps=math.atan2(pi, mu * eps) if (mu * eps or pi) else float("nan"))


def _static_rows(sims) -> List[Dict[str, float]]:
    """dynamical_features.py:27-155 for a list of (same-N) sims, state as is (no kicks)."""
    m = np.stack([s._mass for s in sims])
    q = np.stack([s._pos for s in sims])
    v = np.stack([s._vel for s in sims])
    eps = np.array([math.sqrt(max(float(s.manager.step_s2), 0.0)) for s in sims])
    bk = E.DeviceBucket(m, q, v, eps, sims[0].G, "verlet", sims[0].device)
    bk.prepare(0, 0.0, 0.01, 0.01, 50, want_static=True)
    stat = bk.static.cpu().numpy()
    rows = []
    for i, s in enumerate(sims):
        row = dict(zip(L.STATIC_COLUMNS, (float(x) for x in stat[i])))
        hist = list(s.manager._history)
        row["softening_mean"] = float(np.mean(hist))
        row["softening_std"] = float(np.std(hist))
        rows.append(row)
    return rows


class DynamicalFeatures:
    def __init__(self, sim):
        self.sim = sim
        self.diagnostics = Diagnostics(sim)

    def extract_all(self) -> Dict[str, float]:
        return _static_rows([self.sim])[0]


class EvolutionFeatures:
    def __init__(self, sim, n_samples: int = 20, dt: float = 0.01):
        self.sim, self.n_samples, self.dt = sim, n_samples, dt
        self.diagnostics = Diagnostics(sim)

    def compute_megno(self, n_steps: int, dt: float):
        """evolution_features.py:34-66 -- advances `sim` by n_steps like the reference."""
        sim = self.sim
        n = sim.n_bodies
        raw_r = np.random.randn(n, 2)
        raw_v = np.random.randn(n, 2)
        if _mode_of(sim) == "ham_soft":
            integ = sim._integrator
            b = integ._bucket()
            b.n_sub[:] = max(1, integ.strang_substeps(dt))
            dyn = b.run(dt, 0, 0, n_steps, raw_r[None], raw_v[None], flags=L.RUN_WRITE_STATE, want_dyn=True).cpu().numpy()[0]
            integ._pull(b)
            bk = b.bk
            sim.manager.update_continuous(sim._epsilon)
        else:
            bk = E.DeviceBucket(sim._mass[None], sim._pos[None], sim._vel[None], sim._force_eps(), sim.G, _mode_of(sim), sim.device)
            bk.set_n_sub_from_h(np.array([sim._integrator.h_sub_ref]), dt, sim._integrator.split_n_max)
            dyn = bk.run(dt, 0, 0, n_steps, raw_r[None], raw_v[None], flags=L.RUN_WRITE_STATE).cpu().numpy()[0]
        sim._pos[...] = bk.q.cpu().numpy()[0]
        sim._vel[...] = bk.v.cpu().numpy()[0]
        d = dict(zip(L.DYN_COLUMNS, dyn))
        return float(d["MEGNO"]), float(d["lyapunov_time"])

    def extract_all(self) -> Dict[str, float]:
        feats = DynamicalFeatures(self.sim).extract_all()
        megno, lyap = self.compute_megno(self.n_samples, self.dt)
        feats.update({"MEGNO": megno, "lyapunov_time": lyap, "current_total_energy": self.diagnostics.energy()})
        return feats

    def extract_evolution_features(self):
        f = self.extract_all()
        return {k: f[k] for k in ("MEGNO", "lyapunov_time", "current_total_energy")}


# ---------------------------------------------------------------------------------------------
# batched analysis
# ---------------------------------------------------------------------------------------------

def analyze_simulations(sims, n_steps: int, dt: float, mode: str, via: str = "device") -> List[Dict]:
    """run_stability_analysis (stability_analyzer.py:69-259) + analyze_simulation post-processing
    (batch_stability_analyzer.py:37-58) for a list of NBodySimulation objects, bucketed by (N, mode, G)."""
    n_steps = max(1, int(n_steps))
    interval, n_megno = E.analysis_plan(n_steps, mode)
    rows: List[Dict] = [None] * len(sims)
    # tangent draws in batch order, exactly where the reference draws them (inside each system's analysis)
    draws = {}
    for i, s in enumerate(sims):
        if n_megno > 0 and s.n_bodies > 0:
            draws[i] = (np.random.randn(s.n_bodies, 2), np.random.randn(s.n_bodies, 2))
    buckets: Dict = {}
    for i, s in enumerate(sims):
        if s.n_bodies < 2:
            rows[i] = {"is_stable": float("nan"), "mode": mode}
            continue
        adaptive = bool(getattr(s, "_adaptive_softening", False)) and _mode_of(s) != "ham_soft"
        buckets.setdefault((s.n_bodies, _mode_of(s), float(s.G), str(s.device), adaptive), []).append(i)
    for (N, imode, G, _dev, adaptive), idx in buckets.items():
        group = [sims[i] for i in idx]
        m = np.stack([s._mass for s in group])
        q = np.stack([s._pos for s in group])
        v = np.stack([s._vel for s in group])
        B = len(group)
        rr = np.stack([draws[i][0] for i in idx]) if n_megno > 0 else None
        rv = np.stack([draws[i][1] for i in idx]) if n_megno > 0 else None
        if imode == "ham_soft":
            dyn, stat, status = _analyze_hamsoft(group, m, q, v, G, n_steps, dt, mode, interval, n_megno, rr, rv)
        elif adaptive:
            dyn, stat, status, vk = _analyze_adaptive(group, m, q, v, G, imode, n_steps, dt, mode, interval, n_megno, rr, rv)
            for k, s in enumerate(group):
                s._vel[...] = vk[k]                      # snapshot() mutates the caller's sim
        else:
            eps = np.array([s._force_eps() for s in group])
            # snapshot(): one more corrector half kick with the last |dt| stepped (simulation.py:319-326)
            top = np.array([abs(float(s._integrator._top_dt or s.cfg.initial_dt)) for s in group])
            if mode == "minimal":
                res = _analyze_minimal(m, q, v, eps, G, imode, n_steps, dt, top, group)
                dyn, stat, status, vk = res
            else:
                if np.all(top == top[0]):
                    r = E.analyze_bucket(m, q, v, eps, G, imode, n_steps, dt, mode, rr, rv, L.PREP_SNAPSHOT_KICK,
                                         float(top[0]), float(group[0].cfg.initial_dt), int(group[0].cfg.split_n_max),
                                         group[0].device, via)
                    dyn, stat, status, vk = r.dyn, r.static, r.status, r.v_kicked
                else:   # rare: sims that were last stepped with different dt -> one call per kick size
                    dyn = np.empty((B, L.N_DYN)); stat = np.empty((B, L.N_STATIC)) if mode == "full" else None
                    status = np.zeros(B, dtype=np.int32); vk = np.empty_like(v)
                    for t in np.unique(top):
                        sel = np.where(top == t)[0]
                        r = E.analyze_bucket(m[sel], q[sel], v[sel], eps[sel], G, imode, n_steps, dt, mode,
                                             rr[sel] if rr is not None else None, rv[sel] if rv is not None else None,
                                             L.PREP_SNAPSHOT_KICK, float(t), float(group[0].cfg.initial_dt),
                                             int(group[0].cfg.split_n_max), group[0].device, via)
                        dyn[sel] = r.dyn; status[sel] = r.status; vk[sel] = r.v_kicked
                        if stat is not None:
                            stat[sel] = r.static
            for k, s in enumerate(group):
                s._vel[...] = vk[k]                      # the reference's snapshot() mutates the caller's sim
        for k, i in enumerate(idx):
            s = sims[i]
            d = dict(zip(L.DYN_COLUMNS, (float(x) for x in dyn[k])))
            if mode == "minimal":
                row = {"is_stable": float(d["energy_drift"] < 0.01), "energy_drift": d["energy_drift"], "mode": "minimal"}
            else:
                row = {key: d[key] for key in _DYN_KEYS}
                row["mode"] = mode
                if mode == "full":
                    hist = list(s.manager._history)
                    for c, val in zip(L.STATIC_COLUMNS, stat[k]):
                        row["initial_" + c] = float(val)
                    row["initial_softening_mean"] = float(np.mean(hist))
                    row["initial_softening_std"] = float(np.std(hist))
            if "energy_drift" in row and abs(row["energy_drift"]) > 10:
                print(f"[warning] Extreme energy drift detected: {row['energy_drift']}")
                row["is_stable"] = 0.0
                row["pathological_energy"] = True
            else:
                row["pathological_energy"] = False
            row["softening_policy"] = "adaptive-ham" if imode == "ham_soft" else ("adaptive-classic" if s._adaptive_softening else "static")
            row["_status"] = int(status[k])
            rows[i] = row
    return rows


def _analyze_minimal(m, q, v, eps, G, imode, n_steps, dt, top, group):
    """'minimal' mode: energy drift only.  snapshot() kicks every sim with ITS OWN last |dt|
    (integration_scheme_base.py:154-175), so sims last stepped with different dt are prepared per kick size, like the
    core / full branch."""
    B = m.shape[0]
    dyn = np.empty((B, L.N_DYN)); status = np.zeros(B, dtype=np.int32); vk = np.empty_like(v)
    for t in np.unique(top):
        sel = np.where(top == t)[0]
        bk = E.DeviceBucket(m[sel], q[sel], v[sel], eps[sel], G, imode, group[0].device)
        bk.prepare(L.PREP_SNAPSHOT_KICK, float(t), float(group[0].cfg.initial_dt), dt, int(group[0].cfg.split_n_max))
        vk[sel] = bk.v.cpu().numpy()
        bk.sort()
        dyn[sel] = bk.run(dt, n_steps, 0, 0, flags=L.RUN_ENERGY).cpu().numpy()
        status[sel] = bk.status.cpu().numpy()
    return dyn, None, status, vk


def _analyze_adaptive(group, m, q, v, G, imode, n_steps, dt, mode, interval, n_megno, rr, rv):
    """run_stability_analysis on classic adaptive-softening sims.  snapshot() kicks the original with the CURRENT
    softening; restore() rebuilds an adaptive copy whose manager restarts from the original's `_epsilon` (the constructor
    softening, simulation.py:473-482); the energies use that constant epsilon (diagnostics.py:474)."""
    torch = L.require_cuda()
    B, N = m.shape
    dev = group[0].device
    dyn = np.empty((B, L.N_DYN)); status = np.zeros(B, dtype=np.int32); vk = np.empty_like(v)
    stat = np.empty((B, L.N_STATIC)) if mode == "full" else None
    top = np.array([abs(float(s._integrator._top_dt or s.cfg.initial_dt)) for s in group])
    eps_force = np.array([s._force_eps() for s in group])
    eps_attr = np.array([float(s._epsilon) for s in group])
    hist0 = np.array([list(s.manager._history)[0] for s in group])
    min_soft = np.array([float(s._min_softening) for s in group])
    par = np.stack([np.maximum(hist0, min_soft), min_soft, np.array([float(s._softening_scale) for s in group])], 1)
    for t in np.unique(top):
        sel = np.where(top == t)[0]
        bk = E.DeviceBucket(m[sel], q[sel], v[sel], eps_force[sel], G, imode, dev)
        bk.prepare(L.PREP_SNAPSHOT_KICK, float(t), float(group[0].cfg.initial_dt), dt, int(group[0].cfg.split_n_max),
                   want_static=(mode == "full"))
        vk[sel] = bk.v.cpu().numpy()
        eps_d = E._to_dev(eps_attr[sel], torch.float64, bk.device).clone()
        eps_e = E._to_dev(eps_attr[sel], torch.float64, bk.device)
        par_d = E._to_dev(par[sel], torch.float64, bk.device)
        e_d = torch.zeros((len(sel),), dtype=torch.float64, device=bk.device)
        dyn_d = torch.empty((len(sel), L.N_DYN), dtype=torch.float64, device=bk.device)
        rdr = E._to_dev(rr[sel], torch.float64, bk.device) if n_megno > 0 else None
        rdv = E._to_dev(rv[sel], torch.float64, bk.device) if n_megno > 0 else None
        cfg = group[0].cfg
        with torch.cuda.device(bk.device):
            L.check(L.load().nb_ensemble_analyze_adaptive_f64(
                L.ptr(bk.m), L.ptr(bk.q), L.ptr(bk.v), L.ptr(eps_d), L.ptr(eps_e), L.ptr(par_d), float(G), len(sel), N,
                L.MODES[imode], float(dt), int(n_steps), int(interval if mode != "minimal" else 0), int(n_megno),
                L.ptr(bk.n_sub), L.ptr(rdr), L.ptr(rdv), float(getattr(cfg, "k_wall", 1.0e9)),
                int(getattr(cfg, "barrier_exponent", 5)), L.ptr(e_d), L.ptr(dyn_d), L.ptr(bk.status), L.stream_ptr()),
                "nb_ensemble_analyze_adaptive_f64")
        dyn[sel] = dyn_d.cpu().numpy()
        status[sel] = bk.status.cpu().numpy()
        if stat is not None:
            stat[sel] = bk.static.cpu().numpy()
    return dyn, stat, status, vk


def _analyze_hamsoft(group, m, q, v, G, n_steps, dt, mode, interval, n_megno, rr, rv):
    """restore(snapshot) for ham_soft re-runs the constructor calibration on the snapshotted positions with
    softening = history[0] and then overwrites epsilon, pi, k_soft, mu_soft from the snapshot
    (simulation.py:399-484, simulation_state.py:231-280); no corrector kick (hamiltonian_softening_integrator.py:753)."""
    from . import hamsoft as H
    B = len(group)
    s0 = np.array([list(s.manager._history)[0] for s in group])
    hs = np.concatenate([H.default_params(s.cfg, s0[k], 0.1 * s0[k] if s0[k] > 0 else 0.0)[0] for k, s in enumerate(group)])
    dt0 = float(group[0].cfg.initial_dt)
    b = H.HamSoftBucket(m, q, v, hs, np.stack([np.maximum(s0, hs[:, H.P["eps_min"]]), np.zeros(B)], 1), G, group[0].device)
    b.setup(calibrate=True, freeze_dt=dt0)
    b.hs[:, H.P["k_soft"]] = b.torch.as_tensor(np.array([s._integrator.k_soft for s in group])).to(b.hs.device)
    b.hs[:, H.P["mu_soft"]] = b.torch.as_tensor(np.array([s._integrator.mu_soft for s in group])).to(b.hs.device)
    b.eps_pi[:, 0] = b.torch.as_tensor(np.array([s._epsilon for s in group])).to(b.hs.device)
    b.eps_pi[:, 1] = b.torch.as_tensor(np.array([s._pi for s in group])).to(b.hs.device)
    if abs(abs(dt) - dt0) / dt0 > 0.01:
        b.bump_mu(dt)
        b.setup(calibrate=False, freeze_dt=dt)
    stat = None
    if mode == "full":
        rows = _static_rows(group)
        stat = np.array([[r[c] for c in L.STATIC_COLUMNS] for r in rows])
    dyn = b.run(dt, n_steps, interval if mode != "minimal" else 0, n_megno, rr, rv, flags=L.RUN_ENERGY, want_dyn=True)
    return dyn.cpu().numpy(), stat, b.bk.status.cpu().numpy()


class StabilityAnalyzer:
    def __init__(self, sim, n_steps: int = 1000, dt: float = 0.01, mode: str = "core"):
        self.sim = sim
        self.n_steps = max(1, int(n_steps))
        self.dt = float(dt)
        self.mode = mode
        self.diagnostics = Diagnostics(sim)
        self._initial = (sim._mass.copy(), sim._pos.copy(), sim._vel.copy())      # stability_analyzer.py:42-44

    def run_stability_analysis(self) -> Dict[str, float]:
        row = analyze_simulations([self.sim], self.n_steps, self.dt, self.mode)[0]
        for k in ("pathological_energy", "softening_policy", "_status"):
            row.pop(k, None)
        return row

    def serialize_to_dict(self, diagnostics: Dict[str, float], max_bodies: int = None) -> Dict:
        """stability_analyzer.py:521-561 (its `sim._adaptive` does not exist in the reference -- the call raises
        AttributeError there; here the flag falls back to adaptive_timestep)."""
        sim = self.sim
        m0, q0, v0 = self._initial
        data = {"n_bodies": sim.n_bodies, "G": sim.G, "softening": sim.manager.softening,
                "min_softening": sim._min_softening,
                "adaptive": float(getattr(sim, "_adaptive", getattr(sim, "_adaptive_timestep", False))),
                "integrator_mode": sim._integrator_mode}
        if max_bodies is not None and sim.n_bodies > max_bodies:
            for name, arr in (("mass", m0), ("x", q0[:, 0]), ("y", q0[:, 1]), ("vx", v0[:, 0]), ("vy", v0[:, 1])):
                data[f"{name}_min"] = float(np.min(arr)); data[f"{name}_max"] = float(np.max(arr))
                data[f"{name}_mean"] = float(np.mean(arr)); data[f"{name}_std"] = float(np.std(arr))
        else:
            for i, mass in enumerate(m0):
                data[f"mass_{i}"] = mass
            for i in range(len(q0)):
                data[f"x_{i}"] = q0[i, 0]; data[f"y_{i}"] = q0[i, 1]
            for i in range(len(v0)):
                data[f"vx_{i}"] = v0[i, 0]; data[f"vy_{i}"] = v0[i, 1]
        data.update(diagnostics)
        return data

    def save_to_csv(self, filename: str, diagnostics: Dict[str, float] = None):
        """stability_analyzer.py:563-568."""
        import pandas as pd
        if diagnostics is None:
            diagnostics = self.run_stability_analysis()
        pd.DataFrame([self.serialize_to_dict(diagnostics)]).to_csv(filename, index=False)


class BatchStabilityAnalyzer:
    def __init__(self, n_steps: int = 1000, dt: float = 0.01, mode: str = "core") -> None:
        self.n_steps, self.dt, self.mode = n_steps, dt, mode
        self.results = []

    def analyze_simulation(self, sim) -> Dict[str, float]:
        row = analyze_simulations([sim], self.n_steps, self.dt, self.mode)[0]
        row.pop("_status", None)
        return row

    def analyze_batch(self, simulations, show_progress: bool = True):
        import pandas as pd
        self.results = []
        if show_progress:
            print(f"Analyzing {len(simulations)} simulations...")
        rows = analyze_simulations(list(simulations), self.n_steps, self.dt, self.mode)
        for i, row in enumerate(rows):
            row.pop("_status", None)
            row["simulation_id"] = i
            row["mode"] = self.mode
            self.results.append(row)
        if show_progress:
            print(f"Completed: {len(self.results)} simulations analyzed")
        return pd.DataFrame(self.results)

    def save_batch_results(self, filename: str) -> None:
        import pandas as pd
        if not self.results:
            print("[error] No results to save. Run analyze_batch first.")
            return
        df = pd.DataFrame(self.results)
        df.to_csv(filename, index=False)
        print(f"Saved {len(df)} results to {filename}")

    def get_feature_matrix(self) -> np.ndarray:
        import pandas as pd
        if not self.results:
            print("[error] No results available. Run analyze_batch first.")
            return np.array([])
        return pd.DataFrame(self.results).values
