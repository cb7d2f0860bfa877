"""ham_soft barrier-policy golden vectors (reflection fold / barrier disabled) from the live reference.
Run in the build container:  python oracle/make_golden_hamsoft_policy.py  ->  tests/golden/hamsoft_policies.npz"""
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
from oracle.make_golden_hamsoft import quiet, hamsoft_systems  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    out, names = {}, []
    dt = 0.01
    S = hamsoft_systems(mb)
    for name in ("readme3", "compact_s0.3", "compact3", "compact6"):
        m, p, v, soft, n_steps = S[name]
        for pol, kw in (("reflection", dict(use_soft_barrier=False)), ("disabled", dict(disable_barrier=True)),
                        ("reflection_tight", dict(use_soft_barrier=False))):
            cfg = mb.SimConfig(**kw)
            with quiet():
                sim = mb.NBodySimulation(config=cfg, masses=m, positions=p, velocities=v, softening=soft,
                                         integrator_mode="ham_soft")
            integ = sim._integrator
            if pol == "reflection_tight":
                # a narrow admissible interval around the start value makes epsilon hit both walls within a few steps
                sim._max_softening = float(sim._epsilon) * 1.02
                sim._min_softening = float(sim._epsilon) * 0.995
            key = f"{name}__{pol}_"
            names.append(key)
            out[key + "m"] = m; out[key + "q_in"] = p; out[key + "v_in"] = v; out[key + "soft"] = soft
            out[key + "policy"] = np.array(str(integ.barrier_policy))
            out[key + "flags"] = np.array([float(cfg.use_soft_barrier), float(cfg.disable_barrier)])
            out[key + "ctor"] = np.array([sim._epsilon, sim._pi, sim._min_softening, sim._max_softening,
                                          integ._eps_model._alpha_run, integ.k_soft, integ.mu_soft,
                                          float(integ._frozen_n_sub), integ._omega_spr0])
            out[key + "v0"] = sim._vel.copy()
            out[key + "H0"] = integ.compute_extended_hamiltonian()
            marks = sorted(set([1, max(1, n_steps // 3), n_steps]))
            done = 0
            folds = 0
            prev_pi = sim._pi
            for t in marks:
                for _ in range(t - done):
                    sim.step(dt)
                done = t
                out[key + f"q{t}"] = sim._pos.copy()
                out[key + f"v{t}"] = sim._vel.copy()
                out[key + f"ep{t}"] = np.array([sim._epsilon, sim._pi, integ.mu_soft])
                out[key + f"H{t}"] = integ.compute_extended_hamiltonian()
            out[key + "marks"] = np.array(marks)
            out[key + "n_sub"] = integ._total_substeps_in_last_step
            print(key, "policy", integ.barrier_policy, "eps", sim._epsilon, "in", sim._min_softening, sim._max_softening,
                  "n_sub", integ._total_substeps_in_last_step)
    out["names"] = np.array(names)
    out["dt"] = dt
    np.savez_compressed(os.path.join(OUT, "hamsoft_policies.npz"), **out)


if __name__ == "__main__":
    main()
