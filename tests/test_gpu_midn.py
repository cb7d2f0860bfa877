"""GPU parity for systems of 9 .. 64 bodies (one CTA per system, one body per thread; csrc/ensemble_mid.cu).
The reference accepts any body count in fp64 (simulation.py:39-162); these sizes used to raise here.
Tolerances as for the small-N kernels: 1e-12 per call, 1e-9 after 1000 steps on softened (non-chaotic on this horizon)
systems, feature columns 1e-9 (+ the golden convention for the chaotic ones is not needed: softening keeps them tame)."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def _system(N, seed, scale=3.0):
    rng = np.random.RandomState(seed)
    m = rng.uniform(0.2, 2.0, N)
    q = rng.randn(N, 2) * scale
    v = rng.randn(N, 2) * 0.25
    return m, q, v


def _ring(N, seed, tight_pair=False):
    """Star + N-1 light bodies on near-circular orbits: regular motion (a 1e-13 perturbation grows by < 100 over 1000
    steps, measured with the oracle), so 1000-step parity is meaningful; random N = 12 / 32 clouds amplify by 1e7-1e13."""
    rng = np.random.RandomState(seed)
    m = rng.uniform(0.5, 1.5, N) * 0.02
    m[0] = 1.0
    r = np.concatenate([[0.0], 1.0 + 0.35 * np.arange(N - 1)])
    ph = rng.uniform(0, 2 * np.pi, N)
    q = np.stack([r * np.cos(ph), r * np.sin(ph)], 1)
    vc = np.concatenate([[0.0], np.sqrt(1.0 / r[1:])])
    v = np.stack([-vc * np.sin(ph), vc * np.cos(ph)], 1)
    if tight_pair:
        # the two outermost bodies 0.03 apart: the (unsoftened) schedule asks for several sub-steps per step while the
        # softening (0.15 >> 0.03) keeps their mutual force smooth
        q[N - 1] = q[N - 2] + np.array([0.03, 0.0])
        v[N - 1] = v[N - 2]
    return m, q, v


@pytest.mark.parametrize("N", [9, 12, 32, 64])
def test_pair_and_variational_calls_vs_oracle(N):
    from nbodysimproject_b200 import ensemble as E
    from oracle import nbody_oracle as O
    B = 5
    ms, qs, drs = [], [], []
    for b in range(B):
        m, q, _ = _system(N, 100 + b, scale=1.0 + b)
        ms.append(m); qs.append(q); drs.append(np.random.RandomState(b).randn(N, 2))
    m, q, dr = np.stack(ms), np.stack(qs), np.stack(drs)
    eps = np.array([0.0, 1e-3, 0.05, 0.3, 1e-2])
    acc, U, dV = E.pair_batched(q, m, eps, 1.0)
    da = E.variational_batched(q, m, eps * eps, dr, 1.0)
    acc, U, dV, da = acc.cpu().numpy(), U.cpu().numpy(), dV.cpu().numpy(), da.cpu().numpy()
    for b in range(B):
        assert relerr(acc[b], O.accelerations(q[b], m[b], eps[b], 1.0)) < 1e-12
        assert abs(U[b] - O.softened_potential(q[b], m[b], 1.0, eps[b])) <= 1e-12 * abs(U[b])
        ref = O.dV_d_epsilon(q[b], m[b], eps[b], 1.0)
        assert abs(dV[b] - ref) <= 1e-12 * max(abs(ref), 1e-300)
        assert relerr(da[b], O.variational_accel(q[b], m[b], eps[b] ** 2, dr[b], 1.0)) < 1e-12
        # Newton's third law to rounding
        assert np.max(np.abs((m[b][:, None] * acc[b]).sum(0))) < 1e-13 * np.max(np.abs(m[b][:, None] * acc[b]))


@pytest.mark.parametrize("N,mode,dt", [(12, "verlet", 0.01), (12, "yoshida4", 0.01), (32, "verlet", 0.01),
                                       (32, "yoshida4", 0.01), (12, "yoshida4", 0.1), (24, "verlet", 0.2)])
def test_facade_steps_vs_oracle(N, mode, dt):
    """NBodySimulation(...).step(dt) x 1000 against the oracle: constructor (COM removal, corrector half kick, frozen
    sub-step schedule) and trajectory; the large dt cases run several sub-steps per step."""
    from nbodysimproject_b200 import NBodySimulation
    from oracle import nbody_oracle as O
    m, q, v = _ring(N, N, tight_pair=dt > 0.05)
    sim = NBodySimulation(masses=m, positions=q, velocities=v, softening=0.15, integrator_mode=mode)
    ref = O.OracleSim(m, q, v, softening=0.15, integrator_mode=mode)
    assert sim.n_bodies == N and sim.integrator_mode == mode
    assert abs(sim._integrator.h_sub_ref - ref.h_sub_ref) <= 1e-14 * ref.h_sub_ref
    assert relerr(sim.vel, ref.v) < 1e-13                      # constructor kick
    if dt > 0.05:
        assert ref.n_sub_for(dt) > 1 and sim._integrator.n_sub_for(dt) == ref.n_sub_for(dt)
    done = 0
    marks = ((1, 1e-13), (10, 1e-13), (100, 1e-12), (1000, 1e-9)) if dt < 0.05 else ((1, 1e-13), (10, 1e-12), (100, 1e-10))
    for mark, tol in marks:
        sim.step_many(dt, mark - done)
        for _ in range(mark - done):
            ref.step(dt)
        done = mark
        assert relerr(sim.pos, ref.q) < tol, (mark, relerr(sim.pos, ref.q))
        assert relerr(sim.vel, ref.v) < 10 * tol
    P = (m[:, None] * sim.vel).sum(0)
    assert np.max(np.abs(P)) < 1e-12 * np.max(np.abs(m[:, None] * sim.vel))   # momentum stays at rounding level
    # a short horizon on a random (chaotic) cloud as well: 100 steps
    m, q, v = _system(N, 7 + N)
    sim = NBodySimulation(masses=m, positions=q, velocities=v, softening=0.15, integrator_mode=mode)
    ref = O.OracleSim(m, q, v, softening=0.15, integrator_mode=mode)
    sim.step_many(0.01, 100)
    for _ in range(100):
        ref.step(0.01)
    assert relerr(sim.pos, ref.q) < 1e-10


def test_batch_analysis_features_vs_oracle():
    """run_stability_analysis ('full': sampling, E0/E1 in double-double, MEGNO, 25 static features) for N = 12."""
    from nbodysimproject_b200 import ensemble as E, _lib as L
    from oracle import nbody_oracle as O
    N, B = 12, 3
    rows, ms, qs, vs, rrs, rvs = [], [], [], [], [], []
    for b in range(B):
        m, q, v = _ring(N, 40 + b)
        rng = np.random.RandomState(b)
        rr, rv = rng.randn(N, 2), rng.randn(N, 2)
        sim = O.OracleSim(m, q, v, softening=0.2, integrator_mode="yoshida4")
        rows.append(O.run_stability_analysis(sim, 300, 0.01, "full", rr, rv))
        ms.append(m); qs.append(q); vs.append(v); rrs.append(rr); rvs.append(rv)
    flags = L.PREP_REMOVE_COM | L.PREP_CTOR_KICK | L.PREP_SNAPSHOT_KICK
    for via in ("device", "host"):
        r = E.analyze_bucket(np.stack(ms), np.stack(qs), np.stack(vs), 0.2, 1.0, "yoshida4", 300, 0.01, "full",
                             np.stack(rrs), np.stack(rvs), flags, via=via)
        assert np.all(r.status == 0)
        for b in range(B):
            d = dict(zip(L.DYN_COLUMNS, r.dyn[b]))
            for k in ("com_drift_mean", "com_drift_max", "ang_mom_var_mean", "ang_mom_var_max", "cos_theta_mean", "MEGNO",
                      "lyapunov_time", "is_stable"):
                assert abs(d[k] - rows[b][k]) <= 1e-9 * max(abs(rows[b][k]), 1e-12), (via, b, k, d[k], rows[b][k])
            assert abs(d["energy_drift"] - rows[b]["energy_drift"]) <= 1e-9 * abs(rows[b]["energy_drift"]) + 2e-13
            s = dict(zip(L.STATIC_COLUMNS, r.static[b]))
            for k, val in s.items():
                if k in ("softening_mean", "softening_std"):
                    continue
                ref = rows[b]["initial_" + k]
                assert abs(val - ref) <= 1e-11 * max(abs(ref), 1e-12), (k, val, ref)


def test_whfast_mid_vs_oracle():
    """Star + 9 planets: pseudo-Jacobi Kepler drift (prefix sums on thread 0, Kepler solves one per thread)."""
    from nbodysimproject_b200 import NBodySimulation
    from oracle import nbody_oracle as O
    N = 10
    rng = np.random.RandomState(3)
    m = np.concatenate([[1.0], 10 ** rng.uniform(-6, -4, N - 1)])
    a = 1.0 * 1.6 ** np.arange(N - 1)
    ph = rng.uniform(0, 2 * np.pi, N - 1)
    q = np.concatenate([[[0.0, 0.0]], np.stack([a * np.cos(ph), a * np.sin(ph)], 1)])
    vc = np.sqrt(1.0 / a)
    v = np.concatenate([[[0.0, 0.0]], np.stack([-vc * np.sin(ph), vc * np.cos(ph)], 1)])
    sim = NBodySimulation(masses=m, positions=q, velocities=v, softening=0.0, integrator_mode="whfast")
    ref = O.OracleSim(m, q, v, softening=0.0, integrator_mode="whfast")
    assert sim.integrator_mode == "whfast" and ref.mode == "whfast"
    assert relerr(sim.vel, ref.v) < 1e-13
    sim.step_many(0.05, 200)
    for _ in range(200):
        ref.step(0.05)
    assert relerr(sim.pos, ref.q) < 1e-10 and relerr(sim.vel, ref.v) < 1e-9
    assert sim._status == 0


def test_unsupported_sizes_disable_instead_of_raising(capsys):
    """simulation.py:76-78: the reference never raises; what the kernels do not cover is reported and disabled."""
    from nbodysimproject_b200 import NBodySimulation
    m, q, v = _system(65, 2)
    sim = NBodySimulation(masses=m, positions=q, velocities=v, integrator_mode="ham_soft")
    assert sim.n_bodies == 0 and "ham_soft" in capsys.readouterr().out
    m, q, v = _system(65, 2)
    sim = NBodySimulation(masses=m, positions=q, velocities=v, integrator_mode="verlet")
    assert sim.n_bodies == 0
    sim.step(0.01)
