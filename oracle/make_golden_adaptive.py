"""Classic adaptive-softening golden vectors from the live reference.
Run in the build container:  python oracle/make_golden_adaptive.py  ->  tests/golden/adaptive_softening.npz"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True
sys.modules.setdefault("lightgbm", types.ModuleType("lightgbm"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minbody as mb  # noqa: E402
from oracle.make_golden_hamsoft import quiet  # noqa: E402
from oracle.make_golden import _alt_force  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def systems():
    S = {}
    S["readme3"] = (np.array([1.0, 0.5, 0.1]), np.array([[0, 0], [1, 0], [2, 0.0]]),
                    np.array([[0, 0], [0, 1], [0, 0.5]]), 0.05)
    # an approaching pair: the minimum separation shrinks and re-opens, so epsilon moves both ways and the
    # factor-2 limiter (softening_manager.py:101-104) engages
    S["approach4"] = (np.array([1.0, 0.8, 0.3, 0.2]), np.array([[0, 0], [0.6, 0.05], [2.0, 0.5], [-1.5, 1.0]]),
                      np.array([[0.3, 0], [-0.9, 0.0], [0, 0.4], [0.2, -0.3]]), 0.2)
    gen = mb.InitialConditionGenerator(mb.GeneratorConfig(position_scale=0.8, softening=0.1, seed=3))
    m, p, v = gen.generate_single(5)
    S["random5"] = (m, p, v, 0.1)
    gen = mb.InitialConditionGenerator(mb.GeneratorConfig(position_scale=1.0, softening=0.05, seed=8))
    m, p, v = gen.generate_single(8)
    S["random8"] = (m, p, v, 0.05)
    return S


def main():
    out, names = {}, []
    dt, n_steps = 0.01, 60
    for name, (m, p, v, soft) in systems().items():
        for mode in ("verlet", "yoshida4", "whfast"):
            with quiet():
                sim = mb.NBodySimulation(masses=m, positions=p, velocities=v, softening=soft, integrator_mode=mode,
                                         adaptive_softening=True)
            key = f"{name}__{mode}_"
            names.append(key)
            out[key + "m"] = m; out[key + "q_in"] = p; out[key + "v_in"] = v; out[key + "soft"] = soft
            out[key + "mode_used"] = np.array(str(sim._integrator_mode))
            out[key + "v0"] = sim._vel.copy()
            out[key + "h_sub_ref"] = float(sim._integrator.h_sub_ref)
            # the reference against ITSELF with an equivalent-arithmetic force routine: how far last-bit rounding is
            # amplified by the dynamics and by the kinks of the min-separation rule (the parity horizon)
            import minbody.simulation as simmod
            orig = simmod.gravitational_force
            simmod.gravitational_force = _alt_force
            try:
                with quiet():
                    alt = mb.NBodySimulation(masses=m, positions=p, velocities=v, softening=soft, integrator_mode=mode,
                                             adaptive_softening=True)
            finally:
                simmod.gravitational_force = orig
            eps_t, dE_t, sens_t = [], [], []
            marks = [1, 10, 30, n_steps]
            for t in range(1, n_steps + 1):
                sim.step(dt)
                simmod.gravitational_force = _alt_force
                try:
                    alt.step(dt)
                finally:
                    simmod.gravitational_force = orig
                eps_t.append(sim.manager.s)
                dE_t.append(sim.softening_energy_delta)
                sens_t.append([float(np.max(np.abs(sim._pos - alt._pos)) / np.max(np.abs(sim._pos))),
                               abs(sim.manager.s - alt.manager.s) / abs(sim.manager.s),
                               abs(sim.softening_energy_delta - alt.softening_energy_delta)
                               / max(abs(sim.softening_energy_delta), 1e-300)])
                if t in marks:
                    out[key + f"q{t}"] = sim._pos.copy()
                    out[key + f"v{t}"] = sim._vel.copy()
            out[key + "sens_t"] = np.array(sens_t)
            out[key + "eps_t"] = np.array(eps_t)
            out[key + "dE_t"] = np.array(dE_t)
            out[key + "marks"] = np.array(marks)
            out[key + "n_sub"] = int(sim._integrator._substeps_in_last_step)
            out[key + "history_tail"] = np.array(sim.manager.history[-64:])
            out[key + "history_len"] = len(sim.manager.history)
            print(key, "mode", sim._integrator_mode, "n_sub", sim._integrator._substeps_in_last_step,
                  "eps range", min(eps_t), max(eps_t), "dE", dE_t[-1], "sens(end)", sens_t[-1])
    out["names"] = np.array(names)
    out["dt"] = dt
    np.savez_compressed(os.path.join(OUT, "adaptive_softening.npz"), **out)


if __name__ == "__main__" and "features" not in sys.argv:
    main()


def gen_features():
    """StabilityAnalyzer('full') rows of adaptive-softening simulations (stepped 5 times first, so that manager.s differs
    from the constructor value the restored copy falls back to, simulation.py:473-482), tangent draws recorded, plus the
    reference's own sensitivity under the equivalent-arithmetic force routine.  -> tests/golden/features_adaptive.npz"""
    import contextlib
    import io
    import minbody.simulation as simmod
    out, names = {}, []
    n_steps = 60
    for name, (m, p, v, soft) in systems().items():
        for mode in ("verlet", "yoshida4"):
            rows = []
            draws = []
            for alt in (False, True):
                orig_force = simmod.gravitational_force
                if alt:
                    simmod.gravitational_force = _alt_force
                try:
                    with quiet():
                        sim = mb.NBodySimulation(masses=m, positions=p, velocities=v, softening=soft, integrator_mode=mode,
                                                 adaptive_softening=True)
                        for _ in range(5):
                            sim.step(0.01)
                    orig = np.random.randn
                    if not alt:
                        def rec(*shape):
                            a = orig(*shape)
                            draws.append(a.copy())
                            return a
                        np.random.seed(3)
                        np.random.randn = rec
                    else:
                        it = iter([d.copy() for d in draws])
                        np.random.randn = lambda *shape: next(it)
                    try:
                        with quiet():
                            rows.append(mb.StabilityAnalyzer(sim, n_steps=n_steps, dt=0.01, mode="full").run_stability_analysis())
                    finally:
                        np.random.randn = orig
                finally:
                    simmod.gravitational_force = orig_force
            key = f"{name}__{mode}_"
            names.append(key)
            out[key + "m"] = m; out[key + "q"] = p; out[key + "v"] = v; out[key + "soft"] = soft
            out[key + "raw_r"] = draws[0]; out[key + "raw_v"] = draws[1]
            for c, val in rows[0].items():
                if isinstance(val, (str, bool, np.bool_)):
                    continue
                out[key + "f__" + c] = float(val)
                d = abs(float(val) - float(rows[1][c]))
                out[key + "sens__" + c] = d if np.isfinite(d) else 0.0
            print(key, {k: rows[0][k] for k in ("energy_drift", "MEGNO")}, "sens MEGNO", out[key + "sens__MEGNO"])
    out["names"] = np.array(names)
    out["n_steps"] = n_steps
    out["pre_steps"] = 5
    np.savez_compressed(os.path.join(OUT, "features_adaptive.npz"), **out)


if __name__ == "__main__" and "features" in sys.argv:
    gen_features()
