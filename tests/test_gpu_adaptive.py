"""GPU parity: classic adaptive softening against the live reference's golden vectors
(oracle/make_golden_adaptive.py) through the facade, and the batched C-ABI entry point against the oracle."""
import numpy as np
import pytest

from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu


def test_facade_adaptive_softening_vs_golden():
    import nbodysimproject_b200 as nb
    g = load_golden("adaptive_softening.npz")
    dt = float(g["dt"])
    for key in g["names"]:
        key = str(key)
        mode = key.split("__")[1].rstrip("_")
        sim = nb.NBodySimulation(masses=g[key + "m"], positions=g[key + "q_in"], velocities=g[key + "v_in"],
                                 softening=float(g[key + "soft"]), integrator_mode=mode, adaptive_softening=True)
        assert sim._integrator_mode == str(g[key + "mode_used"])          # whfast + adaptive -> verlet
        assert relerr(sim._vel, g[key + "v0"]) < 1e-14                    # no constructor corrector kick
        assert sim._integrator.h_sub_ref == pytest.approx(float(g[key + "h_sub_ref"]), rel=1e-13)
        eps_t, dE_t = g[key + "eps_t"], g[key + "dE_t"]
        marks = set(int(t) for t in g[key + "marks"])
        # tolerance = base + 100 x the reference's divergence from ITSELF under an equivalent-arithmetic force routine
        # (recorded per step in the golden file): the min-separation rule has kinks that amplify last-bit rounding
        sens = np.maximum.accumulate(g[key + "sens_t"], axis=0)
        for t in range(1, len(eps_t) + 1):
            sim.step(dt)
            sq, se, sd = 100.0 * sens[t - 1]
            assert sim.manager.s == pytest.approx(float(eps_t[t - 1]), rel=1e-9 + se), (key, t)
            assert sim.softening_energy_delta == pytest.approx(float(dE_t[t - 1]), rel=1e-8 + sd, abs=1e-11), (key, t)
            if t in marks:
                assert relerr(sim._pos, g[key + f"q{t}"]) < 1e-9 + sq, (key, t)
                assert relerr(sim._vel, g[key + f"v{t}"]) < 1e-8 + 10 * sq, (key, t)
        assert len(sim.manager.history) == int(g[key + "history_len"])
        assert np.allclose(sim.manager.history[-64:], g[key + "history_tail"], rtol=1e-9 + 100.0 * float(sens[-1, 1]))


def test_step_many_equals_repeated_steps_and_batch_vs_oracle():
    import nbodysimproject_b200 as nb
    from nbodysimproject_b200 import ensemble as E
    from oracle import nbody_oracle as O
    rng = np.random.RandomState(2)
    B, N = 24, 4
    m = rng.uniform(0.3, 2.0, (B, N))
    q = rng.randn(B, N, 2)
    v = rng.randn(B, N, 2) * 0.4
    v -= (m[:, :, None] * v).sum(1, keepdims=True) / m.sum(1)[:, None, None]
    soft = 0.08
    href = np.array([O.classic_h_sub_ref(q[b], m[b], 1.0, 0.01, 50) for b in range(B)])
    qf, vf, hist, dE, st = E.advance_bucket_adaptive(m, q, v, soft, soft, 0.1 * soft, 1.0, href, 1.0, "yoshida4", 0.01, 25)
    assert np.all(st == 0) and hist.shape == (B, 25)
    for b in (0, 5, 23):
        o = O.OracleSim(m[b], q[b], v[b], softening=soft, integrator_mode="yoshida4", adaptive_softening=True,
                        skip_cm_recenter=True)
        eps = []
        for _ in range(25):
            o.step(0.01)
            eps.append(o.s)
        assert np.allclose(hist[b], eps, rtol=1e-9)
        assert relerr(qf[b], o.q) < 1e-9
        assert dE[b] == pytest.approx(o.softening_energy_delta, rel=1e-8, abs=1e-11)
    # facade: one launch of 25 steps == 25 launches of one step
    a = nb.NBodySimulation(masses=m[3], positions=q[3], velocities=v[3], softening=soft, integrator_mode="verlet",
                           adaptive_softening=True)
    a.step_many(0.01, 25)
    b2 = nb.NBodySimulation(masses=m[3], positions=q[3], velocities=v[3], softening=soft, integrator_mode="verlet",
                            adaptive_softening=True)
    for _ in range(25):
        b2.step(0.01)
    assert np.array_equal(a._pos, b2._pos) and a.manager.s == b2.manager.s
    assert a.softening_energy_delta == pytest.approx(b2.softening_energy_delta, rel=1e-12)
    assert a.manager.history == b2.manager.history


@pytest.mark.parametrize("N,mode", [(12, "yoshida4"), (20, "verlet")])
def test_mid_n_adaptive_softening_vs_oracle(N, mode):
    """9 .. 64 bodies (run-time body count, one thread per system; csrc/ensemble_adaptive.cu): stepping through the C ABI
    and through the facade against the oracle, as for N <= 8 above; the reference accepts any body count
    (simulation.py:39-162)."""
    import nbodysimproject_b200 as nb
    from nbodysimproject_b200 import ensemble as E
    from oracle import nbody_oracle as O
    rng = np.random.RandomState(N)
    B = 3
    m = rng.uniform(0.3, 2.0, (B, N))
    q = rng.randn(B, N, 2) * 2.0
    v = rng.randn(B, N, 2) * 0.3
    v -= (m[:, :, None] * v).sum(1, keepdims=True) / m.sum(1)[:, None, None]
    soft = 0.08
    href = np.array([O.classic_h_sub_ref(q[b], m[b], 1.0, 0.01, 50) for b in range(B)])
    qf, vf, hist, dE, st = E.advance_bucket_adaptive(m, q, v, soft, soft, 0.1 * soft, 1.0, href, 1.0, mode, 0.01, 20)
    assert np.all(st == 0) and hist.shape == (B, 20)
    for b in range(B):
        o = O.OracleSim(m[b], q[b], v[b], softening=soft, integrator_mode=mode, adaptive_softening=True,
                        skip_cm_recenter=True)
        eps = []
        for _ in range(20):
            o.step(0.01)
            eps.append(o.s)
        assert np.allclose(hist[b], eps, rtol=1e-9)
        assert relerr(qf[b], o.q) < 1e-9
        assert dE[b] == pytest.approx(o.softening_energy_delta, rel=1e-8, abs=1e-11)
    sim = nb.NBodySimulation(masses=m[0], positions=q[0], velocities=v[0], softening=soft, integrator_mode=mode,
                             adaptive_softening=True)
    o = O.OracleSim(m[0], q[0], v[0], softening=soft, integrator_mode=mode, adaptive_softening=True)
    assert sim.n_bodies == N
    for _ in range(10):
        sim.step(0.01)
        o.step(0.01)
    assert relerr(sim.pos, o.q) < 1e-9
    assert sim.manager.s == pytest.approx(o.s, rel=1e-9)
    assert sim.softening_energy_delta == pytest.approx(o.softening_energy_delta, rel=1e-8, abs=1e-11)


def test_stability_analysis_of_adaptive_sims_vs_golden():
    """StabilityAnalyzer('full') on adaptive-softening sims against the live reference's rows
    (oracle/make_golden_adaptive.py features): tolerance 1e-8 + 100 x the reference's own sensitivity per column."""
    import nbodysimproject_b200 as nb
    g = load_golden("features_adaptive.npz")
    n_steps, pre = int(g["n_steps"]), int(g["pre_steps"])
    cols = ["energy_drift", "angular_momentum_drift", "com_drift_mean", "com_drift_max", "cos_theta_mean", "cos_theta_min",
            "ang_mom_var_mean", "ang_mom_var_max", "MEGNO", "lyapunov_time", "initial_softening_mean",
            "initial_softening_std", "initial_total_energy", "initial_min_separation", "initial_virial_ratio", "is_stable"]
    for key in g["names"]:
        key = str(key)
        mode = key.split("__")[1].rstrip("_")
        sim = nb.NBodySimulation(masses=g[key + "m"], positions=g[key + "q"], velocities=g[key + "v"],
                                 softening=float(g[key + "soft"]), integrator_mode=mode, adaptive_softening=True)
        for _ in range(pre):
            sim.step(0.01)
        draws = [g[key + "raw_r"], g[key + "raw_v"]]
        it = iter(draws)
        orig = np.random.randn
        np.random.randn = lambda *shape: next(it).copy()
        try:
            row = nb.StabilityAnalyzer(sim, n_steps=n_steps, dt=0.01, mode="full").run_stability_analysis()
        finally:
            np.random.randn = orig
        for c in cols:
            ref, sens = float(g[key + "f__" + c]), float(g[key + "sens__" + c])
            if c == "angular_momentum_drift":          # |L1 - L0| / |L0| of ~1e-15: rounding noise on both sides
                assert abs(row[c] - ref) < 1e-12, (key, c, row[c], ref)
                continue
            assert abs(row[c] - ref) <= 1e-8 * max(abs(ref), 1e-12) + 100 * sens, (key, c, row[c], ref)
