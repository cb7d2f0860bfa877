"""GPU parity of ham_soft for systems of 9 .. 64 bodies (one CTA per system, one finite-difference evaluation per thread;
csrc/hamsoft_mid.cu) against the NumPy oracle (oracle/hamsoft_oracle.py, pinned to the live reference by
tests/test_oracle_golden.py).  ham_soft is the reference's default mode (sim_config.py:38) and the reference accepts any
body count (simulation.py:39-162); these sizes used to disable the simulation here.
Tolerances as for the small-N kernels (tests/test_gpu_hamsoft.py): constructor scalars 1e-12, eps* 1e-12, the central
-difference gradient 1e-8 (a difference of two solves divided by 2e-5), H_ext 1e-12, a few Strang sub-steps 1e-8.
The oracle's eps* solve is a Python loop (3.6 ms at N = 9, 0.2 s at N = 64), so the stepping cases keep the sub-step
count small by fixing n_sub on both sides -- the schedule itself is compared first."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def _cluster(N, seed, scale):
    rng = np.random.RandomState(seed)
    m = rng.uniform(0.5, 3.0, N)
    q = rng.randn(N, 2) * scale
    v = rng.randn(N, 2) * 0.4
    v -= (m[:, None] * v).sum(0) / m.sum()
    return m, q, v


def _pair(N, seed, scale, soft, dt0):
    from nbodysimproject_b200 import hamsoft as H
    from nbodysimproject_b200.simulation import SimConfig
    from oracle.hamsoft_oracle import HamSoftOracleSim
    m, q, v = _cluster(N, seed, scale)
    o = HamSoftOracleSim(m, q, v, softening=soft, skip_cm_recenter=True, initial_dt=dt0)
    hs, s0 = H.default_params(SimConfig(), soft, 0.1 * soft)
    b = H.HamSoftBucket(m[None], q[None], v[None], hs, np.array([[s0[0], 0.0]]), 1.0)
    b.setup(calibrate=True, freeze_dt=dt0)
    return o, b


def _ctor_row(b):
    from nbodysimproject_b200.hamsoft import P
    hs = b.hs.cpu().numpy()[0]
    ep = b.eps_pi.cpu().numpy()[0]
    return np.array([ep[0], ep[1], hs[P["eps_min"]], hs[P["eps_max"]], hs[P["alpha_run"]], hs[P["k_soft"]],
                     hs[P["mu_soft"]], float(b.n_sub[0]), hs[P["omega_spr0"]]])


def _oracle_row(o):
    return np.array([o.eps, o.pi, o.eps_min, o.eps_max, o.alpha_run, o.k_soft, o.mu_soft, float(o.frozen_n_sub),
                     o.omega_spr0])


# (N, seed, position scale, softening, dt, sub-steps forced on both sides, macro steps)
CASES = [(9, 9, 0.3, 0.05, 1e-4, 2, 3),       # eps* between the walls: the finite-difference gradient drives the S flow
         (9, 5, 0.3, 0.05, 1e-4, 1, 3),
         (12, 112, 0.5, 0.05, 1e-3, 2, 2),    # eps* clamped at eps_min: analytic fallback + sign alignment
         (20, 120, 0.15, 0.05, 1e-3, 1, 2),
         (33, 133, 1.0, 0.05, 1e-3, 1, 1)]    # 4 N + 1 = 133 evaluations > 128 threads: strided evaluation loop


@pytest.mark.parametrize("N,seed,scale,soft,dt,K,steps", CASES)
def test_mid_hamsoft_constructor_probe_and_steps_vs_oracle(N, seed, scale, soft, dt, K, steps):
    o, b = _pair(N, seed, scale, soft, dt)
    assert np.allclose(_ctor_row(b), _oracle_row(o), rtol=1e-12, atol=0), (_ctor_row(b), _oracle_row(o))
    fb0 = o.n_fallbacks
    es_o, g_o = o.eps_star_and_grad(o.q)
    H_o = o.extended_hamiltonian()
    es, Hx, fb, grad = b.probe()
    assert abs(es[0] - es_o) <= 1e-12 * abs(es_o)
    assert bool(fb[0]) == (o.n_fallbacks > fb0)
    assert np.max(np.abs(grad[0] - g_o)) <= 1e-8 * max(np.max(np.abs(g_o)), 1e-30) + 1e-14, (grad[0], g_o)
    assert abs(Hx[0] - H_o) <= 1e-12 * abs(H_o)
    # a few Strang sub-steps with the same sub-step count on both sides
    o.frozen_n_sub = K
    o.macro_dt_frozen = abs(dt)
    b.n_sub[:] = K
    for _ in range(steps):
        o.step(dt)
    b.run(dt, steps)
    q, v, ep = b.bk.q.cpu().numpy()[0], b.bk.v.cpu().numpy()[0], b.eps_pi.cpu().numpy()[0]
    assert relerr(q, o.q) < 1e-8, relerr(q, o.q)
    assert relerr(v, o.v) < 1e-7, relerr(v, o.v)
    assert abs(ep[0] - o.eps) <= 1e-8 * abs(o.eps), (ep, o.eps)
    assert abs(ep[1] - o.pi) <= 1e-6 * max(abs(o.pi), 1e-9), (ep, o.pi)
    # the forced sub-step count is far below the frozen schedule for the dense clusters, so epsilon may leave
    # [eps_min - R, eps_max + R] (flagged, like the oracle's value); the state must stay finite
    assert int(b.bk.status[0]) & 1 == 0


@pytest.mark.parametrize("use_soft,disabled", [(False, False), (True, True)])
def test_mid_hamsoft_barrier_policies_vs_oracle(use_soft, disabled):
    """Reflection fold (epsilon folded into a squeezed [eps_min, eps_max], pi flipped) and disabled barrier at N = 9
    (hamsoft_utils.py:150-176, hamsoft_stepper.py:72-80, 107-113, 261-303); the oracle's policies are pinned to the live
    reference by tests/test_oracle_golden.py::test_hamsoft_barrier_policies_vs_golden."""
    from nbodysimproject_b200 import hamsoft as H
    from nbodysimproject_b200.hamsoft import P
    from nbodysimproject_b200.simulation import SimConfig
    from oracle.hamsoft_oracle import HamSoftOracleSim
    N, dt, soft, K, steps = 9, 1e-4, 0.05, 2, 4
    m, q, v = _cluster(N, 9, 0.3)
    o = HamSoftOracleSim(m, q, v, softening=soft, skip_cm_recenter=True, initial_dt=dt, use_soft_barrier=use_soft,
                         disable_barrier=disabled)
    cfg = SimConfig()
    cfg.use_soft_barrier, cfg.disable_barrier = use_soft, disabled
    hs, s0 = H.default_params(cfg, soft, 0.1 * soft)
    assert hs[0, P["policy"]] == (2.0 if disabled else 1.0)
    b = H.HamSoftBucket(m[None], q[None], v[None], hs, np.array([[s0[0], 0.0]]), 1.0)
    b.setup(calibrate=True, freeze_dt=dt)
    assert np.allclose(_ctor_row(b), _oracle_row(o), rtol=1e-12, atol=0)
    if not disabled:
        # squeeze the interval around the start value so that the spring flow crosses both walls within a few sub-steps
        lo, hi = o.eps * (1.0 - 2e-4), o.eps * (1.0 + 2e-4)
        o.eps_min, o.eps_max = lo, hi
        b.hs[0, P["eps_min"]] = lo
        b.hs[0, P["eps_max"]] = hi
    o.frozen_n_sub = K
    o.macro_dt_frozen = dt
    b.n_sub[:] = K
    for _ in range(steps):
        o.step(dt)
    b.run(dt, steps)
    ep = b.eps_pi.cpu().numpy()[0]
    assert relerr(b.bk.q.cpu().numpy()[0], o.q) < 1e-8
    assert abs(ep[0] - o.eps) <= 1e-7 * abs(o.eps), (ep, o.eps, o.pi)
    assert abs(ep[1] - o.pi) <= 1e-5 * max(abs(o.pi), 1e-6), (ep, o.eps, o.pi)
    if not disabled:
        assert lo <= ep[0] <= hi


@pytest.mark.parametrize("hook", ["freeze_s_subsystem", "validate_s_only"])
def test_mid_hamsoft_config_hooks_vs_oracle(hook):
    """The two reference test hooks at N = 9 (flags of the parameter row; the oracle's hooks are pinned bit for bit to the
    live reference by tests/test_oracle_golden.py::test_hamsoft_hooks_vs_golden)."""
    from nbodysimproject_b200 import hamsoft as H, _lib as L
    from nbodysimproject_b200.hamsoft import P
    from nbodysimproject_b200.simulation import SimConfig
    from oracle.hamsoft_oracle import HamSoftOracleSim
    N, dt, soft, steps = 9, 1e-4, 0.05, 3
    m, q, v = _cluster(N, 9, 0.3)
    o = HamSoftOracleSim(m, q, v, softening=soft, skip_cm_recenter=True, initial_dt=dt, **{hook: True})
    cfg = SimConfig()
    setattr(cfg, "freeze_s_subsystem" if hook == "freeze_s_subsystem" else "_validate_S_only", True)
    hs, s0 = H.default_params(cfg, soft, 0.1 * soft)
    assert hs[0, P["flags"]] == (L.HS_FLAG_FREEZE_S if hook == "freeze_s_subsystem" else L.HS_FLAG_S_ONLY)
    b = H.HamSoftBucket(m[None], q[None], v[None], hs, np.array([[s0[0], 0.0]]), 1.0)
    b.setup(calibrate=True, freeze_dt=dt)
    assert np.allclose(_ctor_row(b), _oracle_row(o), rtol=1e-12, atol=0)
    if hook == "freeze_s_subsystem":
        o.frozen_n_sub = 2
        o.macro_dt_frozen = dt
        b.n_sub[:] = 2
    for _ in range(steps):
        o.step(dt)
    b.run(dt, steps)
    ep = b.eps_pi.cpu().numpy()[0]
    assert relerr(b.bk.q.cpu().numpy()[0], o.q) < 1e-9
    assert relerr(b.bk.v.cpu().numpy()[0], o.v) < 1e-7
    assert abs(ep[0] - o.eps) <= 1e-8 * abs(o.eps), (ep, o.eps)
    assert abs(ep[1] - o.pi) <= 1e-6 * max(abs(o.pi), 1e-9), (ep, o.pi)


def test_mid_hamsoft_64_bodies_constructor_eps_star_and_invariants():
    """N = 64 (257 evaluations per S half-flow): constructor and eps* against the oracle (the oracle's 257-solve gradient
    takes a minute, so the stepping check is the flow's own invariants: linear momentum is conserved by the pairwise
    forces and by J grad eps* (translation invariance), the run is finite)."""
    N, dt = 64, 1e-3
    o, b = _pair(N, 164, 2.0, 0.05, dt)
    assert np.allclose(_ctor_row(b), _oracle_row(o), rtol=1e-12, atol=0), (_ctor_row(b), _oracle_row(o))
    es, Hx, fb, grad = b.probe()
    es_o = o.eps_target(o.q)
    assert abs(es[0] - es_o) <= 1e-12 * abs(es_o)
    assert abs(Hx[0] - o.extended_hamiltonian()) <= 1e-12 * abs(Hx[0])
    assert np.all(np.isfinite(grad))
    m = b.bk.m.cpu().numpy()[0]
    p0 = (m[:, None] * b.bk.v.cpu().numpy()[0]).sum(0)
    b.n_sub[:] = 2
    b.run(dt, 5)
    v = b.bk.v.cpu().numpy()[0]
    assert np.all(np.isfinite(v)) and np.all(np.isfinite(b.bk.q.cpu().numpy()))
    p1 = (m[:, None] * v).sum(0)
    assert np.max(np.abs(p1 - p0)) < 1e-9 * np.abs(m[:, None] * v).sum()
    assert int(b.bk.status[0]) == 0


def test_mid_hamsoft_batch_is_per_system():
    """Every system of a batch gets the result it gets alone (one CTA per system, no cross-talk), in any launch order."""
    from nbodysimproject_b200 import hamsoft as H
    from nbodysimproject_b200.simulation import SimConfig
    N, B = 10, 7
    rng = np.random.RandomState(3)
    m = rng.uniform(0.5, 3.0, (B, N))
    q = rng.randn(B, N, 2) * rng.uniform(0.2, 1.5, (B, 1, 1))
    v = rng.randn(B, N, 2) * 0.3
    hs, s0 = H.default_params(SimConfig(), 0.05, 0.005, B)
    b = H.HamSoftBucket(m, q, v, hs, np.stack([s0, np.zeros(B)], 1), 1.0)
    b.setup(True, 1e-3)
    b.n_sub.clamp_(max=3)
    ns = b.n_sub.cpu().numpy().copy()
    b.sort()
    b.run(1e-3, 2)
    qa, ea = b.bk.q.cpu().numpy(), b.eps_pi.cpu().numpy()
    for i in (0, 3, 6):
        b1 = H.HamSoftBucket(m[i:i + 1], q[i:i + 1], v[i:i + 1], hs[i:i + 1], np.array([[s0[i], 0.0]]), 1.0)
        b1.setup(True, 1e-3)
        b1.n_sub[:] = int(ns[i])
        b1.run(1e-3, 2)
        assert np.array_equal(b1.bk.q.cpu().numpy()[0], qa[i])
        assert np.array_equal(b1.eps_pi.cpu().numpy()[0], ea[i])


def test_mid_hamsoft_analysis_columns_vs_oracle():
    """The feature row of the ham_soft run kernel at N = 9: H_ext taps, sampled step_metrics, MEGNO with the tangent map
    at the post-step epsilon (diagnostics.py:241-285, 457-549, evolution_features.py:34-66)."""
    from nbodysimproject_b200 import _lib as L
    from oracle import nbody_oracle as O
    N, dt, n_steps, n_megno = 9, 1e-4, 3, 3
    o, b = _pair(N, 9, 0.3, 0.05, dt)
    o.frozen_n_sub = 1
    o.macro_dt_frozen = dt
    b.n_sub[:] = 1
    rng = np.random.RandomState(0)
    rr, rv = rng.randn(1, N, 2), rng.randn(1, N, 2)
    dyn = b.run(dt, n_steps, 1, n_megno, rr, rv, flags=L.RUN_ENERGY | L.RUN_WRITE_STATE, want_dyn=True).cpu().numpy()[0]
    col = {c: i for i, c in enumerate(L.DYN_COLUMNS)}
    E0 = o.extended_hamiltonian()
    L0 = float(O.angular_momentum(o.m, o.q, o.v))
    first, com, var, jeps = None, [], [], []
    for _ in range(n_steps):
        o.step(dt)
        s, first = o.step_metrics(first)
        com.append(s["com_drift"]); var.append(s["var_L"]); jeps.append(s["J_eps"])
    E1 = o.extended_hamiltonian()
    L1 = float(O.angular_momentum(o.m, o.q, o.v))
    dr, dv = O.megno_init_vectors(o.m, rr[0], rv[0])
    t, acc = 0.0, 0.0
    for _ in range(n_megno):
        o.step(dt)
        dr = dr + dv * dt
        dv = dv + O.variational_accel(o.q, o.m, o.eps * o.eps, dr, o.G) * dt
        t += dt
        acc += (np.linalg.norm(dv) / np.linalg.norm(dr)) * t * dt
    megno = 2.0 * acc / t

    def close(a, b_, tol):
        return abs(a - b_) <= tol * max(abs(b_), 1e-300)
    assert close(dyn[col["_E0"]], E0, 1e-12) and close(dyn[col["_E1"]], E1, 1e-11), (dyn[col["_E0"]], E0, dyn[col["_E1"]], E1)
    assert close(dyn[col["_L0"]], L0, 1e-12) and close(dyn[col["_L1"]], L1, 1e-11)
    assert close(dyn[col["com_drift_mean"]], np.mean(com), 1e-8)
    assert close(dyn[col["com_drift_max"]], np.max(com), 1e-8)
    assert close(dyn[col["ang_mom_var_mean"]], np.mean(var), 1e-8)
    assert close(dyn[col["j_eps_mean"]], np.mean(jeps), 1e-6) or abs(dyn[col["j_eps_mean"]] - np.mean(jeps)) < 1e-12
    assert close(dyn[col["MEGNO"]], megno, 1e-7), (dyn[col["MEGNO"]], megno)


def test_facade_default_mode_steps_twelve_bodies():
    """NBodySimulation(...) with 12 bodies in the default mode (ham_soft) steps like the oracle instead of being disabled."""
    import nbodysimproject_b200 as nb
    from oracle.hamsoft_oracle import HamSoftOracleSim
    m, q, v = _cluster(12, 7, 0.8)
    sim = nb.NBodySimulation(masses=m, positions=q, velocities=v, softening=0.05)
    assert sim.n_bodies == 12 and sim.integrator_mode == "ham_soft"
    o = HamSoftOracleSim(m, q, v, softening=0.05)
    assert sim._integrator._frozen_n_sub == o.frozen_n_sub
    dt = 1e-4
    n_o = o.strang_substeps(dt)
    if n_o > 4:                       # keep the oracle's share of the test in seconds
        pytest.skip(f"schedule asks for {n_o} sub-steps")
    for _ in range(2):
        sim.step(dt)
        o.step(dt)
    assert relerr(np.asarray(sim.pos), o.q) < 1e-8
    assert abs(sim._epsilon - o.eps) <= 1e-8 * abs(o.eps)
