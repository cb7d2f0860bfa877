"""Golden numbers for the three self-checks of the reference's validate_ham_soft (hamsoft_validation.py:30-121), taken
from the LIVE reference: (1) extended Hamiltonian before / after 256 steps of dt = 1e-3, (2) the canonical-equation
probe: (epsilon, pi) one step after a snapshot / restore together with the reference's analytic expectations,
(3) the zero-force equilibrium run (G = 0, epsilon = eps*, pi = 0.123456789, 256 steps).
Run in the build container:  python oracle/make_golden_hamsoft_validation.py  ->  tests/golden/hamsoft_validation.npz"""
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
from minbody.diagnostics import Diagnostics  # noqa: E402
from minbody.hamsoft_utils import dU_depsilon_plummer  # noqa: E402
from minbody.barrier import barrier_force  # noqa: E402
from oracle.make_golden_hamsoft import quiet, hamsoft_systems  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    out, names = {}, []
    n_steps, dt = 256, 1e-3
    S = hamsoft_systems(mb)
    for name in ("readme3", "compact3", "compact6"):
        m, p, v, soft, _ = S[name]
        with quiet():
            sim = mb.NBodySimulation(masses=m, positions=p, velocities=v, softening=soft, integrator_mode="ham_soft")
            diag = Diagnostics(sim)
            H0 = diag.compute_extended_hamiltonian()
            for _ in range(n_steps):
                sim.step(dt)
            H1 = diag.compute_extended_hamiltonian()
            snap = sim.snapshot()
            sim_c = mb.NBodySimulation.restore(snap)
            int_c = sim_c._integrator
            eps0, pi0 = float(sim_c._epsilon), float(sim_c._pi)
            eps_star = float(int_c._eps_target(q=sim_c._pos))
            dU = dU_depsilon_plummer(sim_c._pos, sim_c._mass, sim_c.G, eps0)
            Fbar = barrier_force(eps0, float(sim_c._min_softening), float(sim_c._max_softening),
                                 k_wall=float(int_c.k_wall), n=int(int_c._barrier_n()))
            dpi_exp = -(dU + float(int_c.k_soft) * (eps0 - eps_star) - float(Fbar))
            deps_exp = pi0 / float(int_c.mu_soft)
            sim_c.step(dt)
            canon = np.array([eps0, pi0, eps_star, dU, Fbar, float(int_c.mu_soft), dpi_exp, deps_exp,
                              float(sim_c._epsilon), float(sim_c._pi)])
            sim_eq = mb.NBodySimulation.restore(snap)
            sim_eq.G = 0.0
            int_eq = sim_eq._integrator
            sim_eq._epsilon = float(int_eq._eps_target(q=sim_eq._pos))
            sim_eq.manager.update_continuous(sim_eq._epsilon)
            sim_eq._pi = 0.123456789
            eq0 = np.array([sim_eq._epsilon, sim_eq._pi])
            for _ in range(n_steps):
                sim_eq.step(dt)
            eq1 = np.array([sim_eq._epsilon, sim_eq._pi])
        key = name + "_"
        names.append(key)
        out[key + "m"] = m; out[key + "q_in"] = p; out[key + "v_in"] = v; out[key + "soft"] = soft
        out[key + "H"] = np.array([H0, H1])
        out[key + "state256"] = np.concatenate([sim._pos.ravel(), sim._vel.ravel(), [sim._epsilon, sim._pi]])
        out[key + "canon"] = canon
        out[key + "eq"] = np.concatenate([eq0, eq1])
        print(name, "H0,H1", H0, H1, "dH", H1 - H0, "canon dpi num/exp", (canon[9] - pi0) / dt, dpi_exp,
              "deps num/exp", (canon[8] - eps0) / dt, deps_exp, "eq pi drift", eq1[1] - eq0[1])
    out["names"] = np.array(names)
    out["n_steps"] = np.array(n_steps); out["dt"] = np.array(dt)
    np.savez(os.path.join(OUT, "hamsoft_validation.npz"), **out)


if __name__ == "__main__":
    main()
