"""ham_soft test-hook golden vectors from the live reference: SimConfig.freeze_s_subsystem (hamsoft_stepper.py:119-124,
592-600: the S half-flow and the pi half-kick are skipped, epsilon and pi stay frozen) and cfg._validate_S_only
(hamsoft_stepper.py:270-284: a Strang sub-step is S S between the folds, no V and no T).
Run in the build container:  python oracle/make_golden_hamsoft_hooks.py  ->  tests/golden/hamsoft_hooks.npz"""
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
    for name in ("readme3", "compact_s0.3", "compact6"):
        m, p, v, soft, n_steps = S[name]
        for hook in ("freeze_s_subsystem", "_validate_S_only"):
            cfg = mb.SimConfig()
            setattr(cfg, hook, True)
            with quiet():
                sim = mb.NBodySimulation(config=cfg, masses=m, positions=p, velocities=v, softening=soft,
                                         integrator_mode="ham_soft")
            assert bool(getattr(sim.cfg, hook, False)), "the constructor must keep the hook on its config copy"
            integ = sim._integrator
            key = f"{name}__{hook}_"
            names.append(key)
            out[key + "m"] = m; out[key + "q_in"] = p; out[key + "v_in"] = v; out[key + "soft"] = soft
            out[key + "hook"] = np.array(hook)
            out[key + "ctor"] = np.array([sim._epsilon, sim._pi, sim._min_softening, sim._max_softening,
                                          integ._eps_model._alpha_run, integ.k_soft, integ.mu_soft,
                                          float(integ._frozen_n_sub), integ._omega_spr0])
            marks = sorted(set([1, max(1, n_steps // 3), n_steps]))
            done = 0
            for t in marks:
                for _ in range(t - done):
                    sim.step(dt)
                done = t
                out[key + f"q{t}"] = sim._pos.copy()
                out[key + f"v{t}"] = sim._vel.copy()
                out[key + f"ep{t}"] = np.array([sim._epsilon, sim._pi, integ.mu_soft])
            out[key + "marks"] = np.array(marks)
            print(key, "eps", sim._epsilon, "pi", sim._pi, "n_sub", integ._total_substeps_in_last_step,
                  "dq", float(np.abs(sim._pos - p).max()))
    out["names"] = np.array(names)
    out["dt"] = dt
    np.savez_compressed(os.path.join(OUT, "hamsoft_hooks.npz"), **out)


if __name__ == "__main__":
    main()
