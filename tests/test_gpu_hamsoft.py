"""GPU parity for ham_soft (one warp per system; lanes = finite-difference evaluations of eps*).

Tolerances: constructor calibration 1e-12; eps* 1e-12; grad eps* 1e-8 relative to its max (it is a central
difference with step 1e-5: one ulp of eps* is worth 1e-11 in the quotient); trajectories 1e-9 + 100 x the
reference's own sensitivity to an equivalent-arithmetic force routine (recorded in the golden file)."""
import numpy as np
import pytest

from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu


def _bucket(g, key):
    from nbodysimproject_b200 import hamsoft as H, ensemble as E, _lib as L
    from nbodysimproject_b200.simulation import SimConfig
    m, q, v, soft = g[key + "m"], g[key + "q_in"], g[key + "v_in"], float(g[key + "soft"])
    bk0 = E.DeviceBucket(m[None], q[None], v[None], soft, 1.0, "verlet")
    bk0.prepare(L.PREP_REMOVE_COM, 0.0, 0.01, 0.01)                     # simulation.py:85-86
    v0 = bk0.v.cpu().numpy()
    hs, s0 = H.default_params(SimConfig(), soft, 0.1 * soft)
    b = H.HamSoftBucket(m[None], q[None], v0, hs, np.array([[s0[0], 0.0]]), 1.0)
    b.setup(calibrate=True, freeze_dt=0.01)
    return b


def test_hamsoft_constructor_and_probe_vs_golden():
    from nbodysimproject_b200.hamsoft import P
    g = load_golden("hamsoft.npz")
    for key in g["names"]:
        key = str(key)
        b = _bucket(g, key)
        hs = b.hs.cpu().numpy()[0]
        ep = b.eps_pi.cpu().numpy()[0]
        ctor = g[key + "ctor"]   # eps, pi, eps_min, eps_max, alpha_run, k, mu, n_sub, omega0
        mine = np.array([ep[0], ep[1], hs[P["eps_min"]], hs[P["eps_max"]], hs[P["alpha_run"]], hs[P["k_soft"]],
                         hs[P["mu_soft"]], float(b.n_sub[0]), hs[P["omega_spr0"]]])
        assert np.allclose(mine, ctor, rtol=1e-12, atol=0), (key, mine, ctor)
        assert relerr(b.bk.v.cpu().numpy()[0], g[key + "v0"]) < 1e-14
        es, Hx, fb, grad = b.probe()
        assert abs(es[0] - float(g[key + "eps_star0"])) <= 1e-12 * abs(es[0])
        g0 = g[key + "grad0"]
        assert np.max(np.abs(grad[0] - g0)) <= 1e-8 * max(np.max(np.abs(g0)), 1e-30) + 1e-14, (key, grad[0], g0)
        assert abs(Hx[0] - float(g[key + "H0"])) <= 1e-12 * abs(Hx[0])


def test_hamsoft_trajectories_vs_golden():
    import nbodysimproject_b200._lib as L
    g = load_golden("hamsoft.npz")
    dt = float(g["dt"])
    for key in g["names"]:
        key = str(key)
        b = _bucket(g, key)
        assert int(b.n_sub[0]) == int(g[key + "n_sub"])
        done = 0
        for mark in g[key + "marks"]:
            mark = int(mark)
            b.run(dt, mark - done)
            done = mark
            sens = g[key + f"sens{mark}"]
            tol = 1e-9 + 100.0 * sens[0]
            assert relerr(b.bk.q.cpu().numpy()[0], g[key + f"q{mark}"]) < tol, (key, mark)
            assert relerr(b.bk.v.cpu().numpy()[0], g[key + f"v{mark}"]) < 10 * tol, (key, mark)
            ep = b.eps_pi.cpu().numpy()[0]
            ref = g[key + f"ep{mark}"]
            assert abs(ep[0] - ref[0]) <= (1e-9 + 100 * sens[1]) * abs(ref[0]), (key, mark, ep, ref)
            assert abs(ep[1] - ref[1]) <= (1e-7 + 100 * sens[2]) * max(abs(ref[1]), 1e-9), (key, mark, ep, ref)
            assert abs(float(b.hs[0, 1]) - ref[2]) <= 1e-12 * ref[2]
            _, Hx, _, _ = b.probe()
            Href = float(g[key + f"H{mark}"])
            assert abs(Hx[0] - Href) <= 1e-6 * max(abs(Href), 1.0), (key, mark, Hx[0], Href)
        assert int(b.bk.status[0]) == 0


def test_hamsoft_batch_matches_single_and_oracle():
    """A mixed batch (one warp per system) gives each system the result it gets alone; spot-check the oracle."""
    from nbodysimproject_b200 import hamsoft as H, ensemble as E, _lib as L
    from nbodysimproject_b200.simulation import SimConfig
    from oracle.hamsoft_oracle import HamSoftOracleSim
    rng = np.random.RandomState(4)
    B, N = 37, 4
    m = rng.uniform(0.5, 3.0, (B, N))
    q = rng.randn(B, N, 2) * rng.uniform(0.2, 1.0, (B, 1, 1))
    v = rng.randn(B, N, 2) * 0.4
    v -= (m[:, :, None] * v).sum(1, keepdims=True) / m.sum(1)[:, None, None]
    hs, s0 = H.default_params(SimConfig(), 0.05, 0.005, B)
    b = H.HamSoftBucket(m, q, v, hs, np.stack([s0, np.zeros(B)], 1), 1.0)
    b.setup(True, 0.01)
    nsub = b.n_sub.cpu().numpy().copy()
    b.run(0.01, 3)
    qa, ea = b.bk.q.cpu().numpy(), b.eps_pi.cpu().numpy()
    for i in (0, 7, 36):
        b1 = H.HamSoftBucket(m[i:i + 1], q[i:i + 1], v[i:i + 1], hs[i:i + 1], np.array([[s0[i], 0.0]]), 1.0)
        b1.setup(True, 0.01)
        b1.run(0.01, 3)
        assert np.array_equal(b1.bk.q.cpu().numpy()[0], qa[i])
        assert np.array_equal(b1.eps_pi.cpu().numpy()[0], ea[i])
        o = HamSoftOracleSim(m[i], q[i], v[i], softening=0.05, skip_cm_recenter=True)
        assert o.frozen_n_sub == nsub[i]
        for _ in range(3):
            o.step(0.01)
        assert relerr(qa[i], o.q) < 1e-8
        assert abs(ea[i, 0] - o.eps) < 1e-8 * abs(o.eps)


def test_hamsoft_barrier_policies_vs_golden():
    """Reflection fold and disabled barrier (hamsoft_utils.py:150-176, hamsoft_stepper.py:72-80, 107-113, 261-303) against
    the live reference's outputs; the 'tight' cases fold epsilon at both walls of a squeezed [eps_min, eps_max]."""
    from nbodysimproject_b200 import hamsoft as H, ensemble as E, _lib as L
    from nbodysimproject_b200.simulation import SimConfig
    from nbodysimproject_b200.hamsoft import P
    g = load_golden("hamsoft_policies.npz")
    dt = float(g["dt"])
    for key in g["names"]:
        key = str(key)
        use_soft, disabled = [bool(x) for x in g[key + "flags"]]
        m, q, v, soft = g[key + "m"], g[key + "q_in"], g[key + "v_in"], float(g[key + "soft"])
        bk0 = E.DeviceBucket(m[None], q[None], v[None], soft, 1.0, "verlet")
        bk0.prepare(L.PREP_REMOVE_COM, 0.0, 0.01, 0.01)
        cfg = SimConfig()
        cfg.use_soft_barrier, cfg.disable_barrier = use_soft, disabled
        hs, s0 = H.default_params(cfg, soft, 0.1 * soft)
        assert hs[0, P["policy"]] == (2.0 if disabled else 1.0)
        b = H.HamSoftBucket(m[None], q[None], bk0.v.cpu().numpy(), hs, np.array([[s0[0], 0.0]]), 1.0)
        b.setup(calibrate=True, freeze_dt=0.01)
        ctor = g[key + "ctor"]
        if "tight" in key:
            b.hs[0, P["eps_min"]] = float(ctor[2])
            b.hs[0, P["eps_max"]] = float(ctor[3])
        hsv, ep = b.hs.cpu().numpy()[0], b.eps_pi.cpu().numpy()[0]
        mine = np.array([ep[0], ep[1], hsv[P["eps_min"]], hsv[P["eps_max"]], hsv[P["alpha_run"]], hsv[P["k_soft"]],
                         hsv[P["mu_soft"]], float(b.n_sub[0]), hsv[P["omega_spr0"]]])
        assert np.allclose(mine, ctor, rtol=1e-12, atol=0), (key, mine, ctor)
        done = 0
        for mark in g[key + "marks"]:
            mark = int(mark)
            b.run(dt, mark - done)
            done = mark
            ep = b.eps_pi.cpu().numpy()[0]
            ref = g[key + f"ep{mark}"]
            assert relerr(b.bk.q.cpu().numpy()[0], g[key + f"q{mark}"]) < 1e-8, (key, mark)
            assert abs(ep[0] - ref[0]) <= 1e-7 * abs(ref[0]), (key, mark, ep, ref)
            assert abs(ep[1] - ref[1]) <= 1e-5 * max(abs(ref[1]), 1e-6), (key, mark, ep, ref)
            if not disabled:
                assert hsv[P["eps_min"]] <= ep[0] <= hsv[P["eps_max"]]


def test_hamsoft_two_systems_per_warp_for_small_n():
    """N <= 4: the finite-difference evaluations fit a half warp (N = 4 with the unperturbed solve done cooperatively),
    so two systems share a warp (half-warp shuffles, the pair runs to the larger sub-step count).  An odd batch with
    mixed n_sub must give every system exactly what it gets alone."""
    from nbodysimproject_b200 import hamsoft as H
    from nbodysimproject_b200.simulation import SimConfig
    rng = np.random.RandomState(11)
    for N in (2, 3, 4):
        B = 7
        m = rng.uniform(0.5, 3.0, (B, N))
        q = rng.randn(B, N, 2) * rng.uniform(0.15, 1.0, (B, 1, 1))
        v = rng.randn(B, N, 2) * 0.4
        v -= (m[:, :, None] * v).sum(1, keepdims=True) / m.sum(1)[:, None, None]
        hs, s0 = H.default_params(SimConfig(), 0.05, 0.005, B)
        b = H.HamSoftBucket(m, q, v, hs, np.stack([s0, np.zeros(B)], 1), 1.0)
        b.setup(True, 0.01)
        nsub = b.n_sub.cpu().numpy().copy()
        rr, rv = rng.randn(B, N, 2), rng.randn(B, N, 2)
        dyn = b.run(0.01, 6, 2, 3, rr, rv, flags=3, want_dyn=True).cpu().numpy()
        qa, ea = b.bk.q.cpu().numpy(), b.eps_pi.cpu().numpy()
        if N >= 3:
            assert len(np.unique(nsub)) > 1                    # the pairs really mix sub-step counts
        for i in range(B):
            b1 = H.HamSoftBucket(m[i:i + 1], q[i:i + 1], v[i:i + 1], hs[i:i + 1], np.array([[s0[i], 0.0]]), 1.0)
            b1.setup(True, 0.01)
            d1 = b1.run(0.01, 6, 2, 3, rr[i:i + 1], rv[i:i + 1], flags=3, want_dyn=True).cpu().numpy()
            assert np.array_equal(b1.bk.q.cpu().numpy()[0], qa[i]), (N, i)
            assert np.array_equal(b1.eps_pi.cpu().numpy()[0], ea[i]), (N, i)
            assert np.array_equal(d1[0], dyn[i], equal_nan=True), (N, i)
