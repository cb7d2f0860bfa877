"""ham_soft golden vectors from the live reference (called by oracle/make_golden.py hamsoft)."""
from __future__ import annotations

import contextlib
import io
import os

import numpy as np


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def hamsoft_systems(mb):
    S = {}
    S["readme3"] = (np.array([1.0, 0.5, 0.1]), np.array([[0, 0], [1, 0], [2, 0.0]]),
                    np.array([[0, 0], [0, 1], [0, 0.5]]), 1e-3, 100)
    # compact systems exercise the eps* gradient / J-cap path
    for s, steps in ((1.0, 30), (0.3, 30), (0.1, 12)):
        gen = mb.InitialConditionGenerator(mb.GeneratorConfig(position_scale=s, softening=0.05, seed=5))
        m, p, v = gen.generate_single(4)
        S[f"compact_s{s}"] = (m, p, v, 0.05, steps)
    gen = mb.InitialConditionGenerator(mb.GeneratorConfig(position_scale=0.5, softening=0.05, seed=11))
    m, p, v = gen.generate_single(3)
    S["compact3"] = (m, p, v, 0.05, 30)
    gen = mb.InitialConditionGenerator(mb.GeneratorConfig(position_scale=0.6, softening=0.08, seed=23))
    m, p, v = gen.generate_single(6)
    S["compact6"] = (m, p, v, 0.08, 10)
    return S


def _alt_force(q, m, eps=0.0, G=1.0):
    q = np.asarray(q, dtype=float)
    m = np.asarray(m, dtype=float)
    F = np.zeros_like(q)
    for i in range(len(m)):
        for j in range(i + 1, len(m)):
            d = q[i] - q[j]
            r2 = d[0] * d[0] + d[1] * d[1] + eps * eps
            if r2 > 0.0:
                w = 1.0 / np.sqrt(r2)
                f = G * m[i] * m[j] * w * w * w * d
                F[i] -= f
                F[j] += f
    return F


def gen_hamsoft(mb, OUT):
    import minbody.hamsoft_stepper as stepper_mod
    out = {}
    names = []
    dt = 0.01
    for name, (m, p, v, soft, n_steps) in hamsoft_systems(mb).items():
        with quiet():
            sim = mb.NBodySimulation(masses=m, positions=p, velocities=v, softening=soft, integrator_mode="ham_soft")
        integ = sim._integrator
        key = name + "_"
        names.append(key)
        out[key + "m"] = m; out[key + "q_in"] = p; out[key + "v_in"] = v; out[key + "soft"] = soft
        out[key + "n_steps"] = n_steps
        out[key + "ctor"] = np.array([sim._epsilon, sim._pi, sim._min_softening, sim._max_softening,
                                      integ._eps_model._alpha_run, integ.k_soft, integ.mu_soft,
                                      float(integ._frozen_n_sub), integ._omega_spr0])
        out[key + "v0"] = sim._vel.copy()
        es, g = integ.eps_star_and_grad(sim._pos)
        out[key + "eps_star0"] = es
        out[key + "grad0"] = g
        out[key + "H0"] = integ.compute_extended_hamiltonian()
        # one S half-flow tap on a copy of the state
        with quiet():
            probe = mb.NBodySimulation(masses=m, positions=p, velocities=v, softening=soft, integrator_mode="ham_soft")
        h = dt / float(probe._integrator._frozen_n_sub)
        probe._integrator._hs_stepper.s_half(h)
        info = probe._integrator._last_s_info
        out[key + "tap_s"] = np.array([info["I_tau"], info["J"], info["J_applied"], info["eps_star"], info["theta"],
                                       info["barrier_kick1"], info["barrier_kick2"], probe._epsilon, probe._pi])
        out[key + "tap_s_v"] = probe._vel.copy()
        probe._integrator._hs_stepper.v_half_kick(h, eps_override=float(probe._epsilon))
        vk = probe._integrator._last_vkick
        out[key + "tap_v"] = np.array([vk["dVgrav_deps"], vk["dSbar_deps"], probe._pi])
        out[key + "tap_v_v"] = probe._vel.copy()
        # trajectory + sensitivity (reference vs itself with an equivalent-arithmetic force routine)
        with quiet():
            alt = mb.NBodySimulation(masses=m, positions=p, velocities=v, softening=soft, integrator_mode="ham_soft")
        orig = stepper_mod._grav_force
        marks = sorted(set([1, max(1, n_steps // 3), n_steps]))
        done = 0
        for t in marks:
            for _ in range(t - done):
                sim.step(dt)
                stepper_mod._grav_force = _alt_force
                try:
                    alt.step(dt)
                finally:
                    stepper_mod._grav_force = orig
            done = t
            out[key + f"q{t}"] = sim._pos.copy()
            out[key + f"v{t}"] = sim._vel.copy()
            out[key + f"ep{t}"] = np.array([sim._epsilon, sim._pi, integ.mu_soft])
            out[key + f"H{t}"] = integ.compute_extended_hamiltonian()
            scale = max(float(np.max(np.abs(sim._pos))), 1e-300)
            out[key + f"sens{t}"] = np.array([float(np.max(np.abs(sim._pos - alt._pos))) / scale,
                                              abs(sim._epsilon - alt._epsilon) / abs(sim._epsilon),
                                              abs(sim._pi - alt._pi) / max(abs(sim._pi), 1e-300)])
        out[key + "marks"] = np.array(marks)
        out[key + "n_sub"] = integ._total_substeps_in_last_step
    out["names"] = np.array(names)
    out["dt"] = dt
    np.savez_compressed(os.path.join(OUT, "hamsoft.npz"), **out)
    print("hamsoft:", len(names))
    gen_hamsoft_features(mb, OUT)


def gen_hamsoft_features(mb, OUT):
    """BatchStabilityAnalyzer('full') rows for ham_soft sims (default integrator mode), tangent draws recorded,
    plus the per-column sensitivity under the equivalent-arithmetic force routine."""
    import minbody.hamsoft_stepper as stepper_mod
    S = hamsoft_systems(mb)
    keys = ["readme3", "compact_s1.0", "compact3", "compact6"]
    n_steps = 40

    def build():
        sims = []
        for k in keys:
            m, p, v, soft, _ = S[k]
            with quiet():
                sims.append(mb.NBodySimulation(masses=m, positions=p, velocities=v, softening=soft))
        return sims

    draws = []
    orig = np.random.randn

    def rec(*shape):
        a = orig(*shape)
        draws.append(a.copy())
        return a

    np.random.seed(99)
    np.random.randn = rec
    try:
        with quiet():
            df = mb.BatchStabilityAnalyzer(n_steps=n_steps, dt=0.01, mode="full").analyze_batch(build(), show_progress=False)
    finally:
        np.random.randn = orig
    it = iter([d.copy() for d in draws])
    np.random.randn = lambda *shape: next(it)
    of = stepper_mod._grav_force
    stepper_mod._grav_force = _alt_force
    try:
        with quiet():
            df2 = mb.BatchStabilityAnalyzer(n_steps=n_steps, dt=0.01, mode="full").analyze_batch(build(), show_progress=False)
    finally:
        np.random.randn = orig
        stepper_mod._grav_force = of
    out = {"names": np.array(keys), "n_steps": n_steps, "dt": 0.01, "columns": np.array(list(df.columns)), "seed": 99}
    for i, k in enumerate(keys):
        m, p, v, soft, _ = S[k]
        out[f"{k}_m"] = m; out[f"{k}_q"] = p; out[f"{k}_v"] = v; out[f"{k}_soft"] = soft
        out[f"{k}_raw_r"] = draws[2 * i]; out[f"{k}_raw_v"] = draws[2 * i + 1]
        for c in df.columns:
            val = df.iloc[i][c]
            if isinstance(val, (str, bool, np.bool_)):
                out[f"{k}__{c}"] = np.array(str(val))
            else:
                out[f"{k}__{c}"] = float(val)
                d = abs(float(val) - float(df2.iloc[i][c]))
                out[f"{k}__sens__{c}"] = d if np.isfinite(d) else 0.0
    np.savez_compressed(os.path.join(OUT, "features_ham_soft.npz"), **out)
    print("features ham_soft", df.shape)
