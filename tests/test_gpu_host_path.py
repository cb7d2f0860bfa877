"""GPU: the HOST entry point nb_ensemble_analyze_host_ex (BatchStabilityAnalyzer.analyze_batch for one bucket as one
C-ABI call, batch_stability_analyzer.py:62-80): chunk pipelining is invisible in the results, the option flags do what
the header says, and the reference's DEFAULT mode (ham_soft, sim_config.py:38) and classic adaptive softening go through
it with the same numbers as the device-pointer path (which the golden tests pin against the reference)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bucket(B, N, seed=3):
    rng = np.random.RandomState(seed)
    m = rng.uniform(0.5, 3.0, (B, N))
    q = rng.randn(B, N, 2) * rng.uniform(0.3, 1.5, (B, 1, 1))
    v = rng.randn(B, N, 2) * 0.4
    eps = rng.uniform(0.01, 0.08, B)
    rr, rv = rng.randn(B, N, 2), rng.randn(B, N, 2)
    return m, q, v, eps, rr, rv


def test_chunked_equals_unchunked_and_device_path():
    from nbodysimproject_b200 import ensemble as E, _lib as L
    for N, mode in ((3, "yoshida4"), (6, "verlet"), (4, "whfast")):
        B = 5003
        m, q, v, eps, rr, rv = _bucket(B, N)
        if mode == "whfast":
            eps = np.zeros(B)
            m[:, 0] = 50.0
        flags = L.PREP_REMOVE_COM | L.PREP_CTOR_KICK | L.PREP_SNAPSHOT_KICK
        one = E.analyze_host(m, q, v.copy(), eps, 1.0, mode, 120, 0.01, 30, rr, rv, flags, n_chunks=1)
        many = E.analyze_host(m, q, v.copy(), eps, 1.0, mode, 120, 0.01, 30, rr, rv, flags, n_chunks=7)
        dev = E.analyze_bucket(m, q, v.copy(), eps, 1.0, mode, 120, 0.01, "full", rr, rv, flags, via="device")
        # analysis_plan("full") with 120 steps gives the same 30 MEGNO steps? keep the device run explicit instead
        for a, b in ((one, many),):
            assert np.array_equal(a.dyn, b.dyn, equal_nan=True), (N, mode)
            assert np.array_equal(a.static, b.static, equal_nan=True)
            assert np.array_equal(a.n_sub, b.n_sub) and np.array_equal(a.status, b.status)
            assert np.array_equal(a.v_kicked, b.v_kicked)
        assert np.array_equal(one.static, dev.static, equal_nan=True)
        assert np.array_equal(one.v_kicked, dev.v_kicked)
        cols = [L.DYN_COLUMNS.index(c) for c in ("energy_drift", "angular_momentum_drift", "com_drift_mean")]
        assert np.array_equal(one.dyn[:, cols], dev.dyn[:, cols], equal_nan=True)


def test_host_flags_compact_keep_v_device_tangent():
    from nbodysimproject_b200 import ensemble as E, _lib as L
    B, N = 3001, 5
    m, q, v, eps, rr, rv = _bucket(B, N, seed=9)
    flags = L.PREP_REMOVE_COM | L.PREP_CTOR_KICK
    full = E.analyze_host(m, q, v.copy(), eps, 1.0, "yoshida4", 100, 0.01, 20, rr, rv, flags)
    v_in = v.copy()
    comp = E.analyze_host(m, q, v_in, eps, 1.0, "yoshida4", 100, 0.01, 20, rr, rv, flags, compact=True, keep_v=True,
                          n_chunks=3)
    assert comp.dyn.shape == (B, L.N_DYN_USER)
    assert np.array_equal(comp.dyn, full.dyn[:, :L.N_DYN_USER], equal_nan=True)
    assert np.array_equal(v_in, v)                       # NB_HOST_KEEP_V: caller's velocities untouched
    assert not np.array_equal(full.v_kicked, v)          # default: mutated like the reference's snapshot()
    # device-drawn tangents: a function of (seed, global index) only -> a shard sees the draws of the full ensemble
    t_all = E.analyze_host(m, q, v.copy(), eps, 1.0, "yoshida4", 100, 0.01, 20, None, None, flags, tangent_seed=77,
                           first_index=1000)
    lo, hi = 1200, 2500
    t_sh = E.analyze_host(m[lo:hi], q[lo:hi], v[lo:hi].copy(), eps[lo:hi], 1.0, "yoshida4", 100, 0.01, 20, None, None,
                          flags, tangent_seed=77, first_index=1000 + lo, n_chunks=2)
    assert np.array_equal(t_all.dyn[lo:hi], t_sh.dyn, equal_nan=True)
    ci = L.DYN_COLUMNS.index("MEGNO")
    assert np.all(np.isfinite(t_all.dyn[:, ci])) and np.std(t_all.dyn[:, ci]) > 0
    assert not np.array_equal(t_all.dyn[:, ci], full.dyn[:, ci])
    # everything that does not depend on the tangent draws is unchanged
    ce = L.DYN_COLUMNS.index("energy_drift")
    assert np.array_equal(t_all.dyn[:, ce], full.dyn[:, ce], equal_nan=True)


def test_device_tangent_draws_are_standard_normal():
    import torch
    from nbodysimproject_b200 import _lib as L
    B, N = 20000, 4
    dr = torch.empty((B, N, 2), dtype=torch.float64, device="cuda")
    dv = torch.empty_like(dr)
    L.check(L.load().nb_generate_tangent_f64(N, B, 5, 0, L.ptr(dr), L.ptr(dv), L.stream_ptr()), "tangent")
    x = torch.cat([dr.flatten(), dv.flatten()]).cpu().numpy()
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1.0) < 0.01
    assert abs(np.corrcoef(dr.flatten().cpu().numpy(), dv.flatten().cpu().numpy())[0, 1]) < 0.01


def test_hamsoft_through_host_entry_matches_device_path():
    """The reference's default mode through the host C entry point: same kernels, same order as the Python-driven
    sequence the golden tests pin (tests/test_gpu_hamsoft.py), so the numbers are identical."""
    from nbodysimproject_b200 import ensemble as E, hamsoft as H, _lib as L
    from nbodysimproject_b200.simulation import SimConfig
    for N in (3, 5, 10):              # 10: one CTA per system (csrc/hamsoft_mid.cu)
        B = 257 if N <= 8 else 17
        m, q, v, eps, rr, rv = _bucket(B, N, seed=21)
        if N > 8:
            # dense 10-body clusters clamp eps* at eps_min, and the frozen pi budget of the reference then asks for up to
            # 1.6e5 sub-steps per step (hamiltonian_softening_integrator.py:1125-1221: no cap); 4x wider systems ask for <= 34
            q = q * 4.0
        soft = np.full(B, 0.05)
        r = E.analyze_host(m, q, v.copy(), soft, 1.0, "ham_soft", 40, 0.01, 10, rr, rv, L.PREP_REMOVE_COM, n_chunks=3,
                           eps_pi=np.stack([soft, np.zeros(B)], 1))
        # device-pointer sequence
        bk0 = E.DeviceBucket(m, q, v.copy(), soft, 1.0, "verlet")
        bk0.prepare(L.PREP_REMOVE_COM, 0.0, 0.01, 0.01)
        v0 = bk0.v.cpu().numpy()
        hs, s0 = H.default_params(SimConfig(), soft, 0.1 * soft, B)
        b = H.HamSoftBucket(m, q, v0, hs, np.stack([s0, np.zeros(B)], 1), 1.0)
        b.setup(calibrate=True, freeze_dt=0.01)
        dyn = b.run(0.01, 40, 1, 10, rr, rv, flags=L.RUN_ENERGY | L.RUN_WRITE_STATE, want_dyn=True).cpu().numpy()
        assert np.array_equal(r.n_sub, b.n_sub.cpu().numpy())
        assert N <= 8 or int(r.n_sub.max()) <= 34
        assert np.array_equal(r.dyn, dyn, equal_nan=True), N
        assert np.array_equal(r.extra["eps_pi"], b.eps_pi.cpu().numpy(), equal_nan=True)
        assert np.array_equal(r.v_kicked, v0)
        assert np.all(np.isfinite(r.static))
        assert np.mean(r.status == 0) > 0.8
        # default start state (no eps_pi given) is the same thing
        r2 = E.analyze_host(m, q, v.copy(), soft, 1.0, "ham_soft", 40, 0.01, 10, rr, rv, L.PREP_REMOVE_COM)
        assert np.array_equal(r2.dyn, r.dyn, equal_nan=True)


def test_adaptive_through_host_entry_matches_device_path():
    import torch
    from nbodysimproject_b200 import ensemble as E, _lib as L
    B, N = 300, 4
    m, q, v, eps, rr, rv = _bucket(B, N, seed=5)
    soft = np.full(B, 0.05)
    par = np.stack([soft, 0.1 * soft, np.full(B, 0.05)], 1)
    r = E.analyze_host(m, q, v.copy(), soft, 1.0, "verlet", 60, 0.01, 10, rr, rv, L.PREP_SNAPSHOT_KICK, soft_par=par,
                       n_chunks=2)
    bk = E.DeviceBucket(m, q, v.copy(), soft, 1.0, "verlet")
    bk.prepare(L.PREP_SNAPSHOT_KICK, 0.01, 0.01, 0.01, 50, want_static=True)
    dev = bk.device
    eps_d = torch.as_tensor(soft).to(dev).clone()
    eps_e = torch.as_tensor(soft).to(dev)
    par_d = torch.as_tensor(par).to(dev)
    e_d = torch.zeros((B,), dtype=torch.float64, device=dev)
    dyn_d = torch.empty((B, L.N_DYN), dtype=torch.float64, device=dev)
    rr_d, rv_d = torch.as_tensor(rr).to(dev), torch.as_tensor(rv).to(dev)      # named: they must outlive the launch
    L.check(L.load().nb_ensemble_analyze_adaptive_f64(
        L.ptr(bk.m), L.ptr(bk.q), L.ptr(bk.v), L.ptr(eps_d), L.ptr(eps_e), L.ptr(par_d), 1.0, B, N, L.MODES["verlet"], 0.01,
        60, 1, 10, L.ptr(bk.n_sub), L.ptr(rr_d), L.ptr(rv_d), 1.0e9, 5,
        L.ptr(e_d), L.ptr(dyn_d), L.ptr(bk.status), L.stream_ptr()), "adaptive")
    assert np.array_equal(r.dyn, dyn_d.cpu().numpy(), equal_nan=True)
    assert np.array_equal(r.extra["energy_delta"], e_d.cpu().numpy(), equal_nan=True)


def test_host_entry_leaves_current_device_and_rejects_bad_input():
    import ctypes
    import torch
    from nbodysimproject_b200 import ensemble as E, _lib as L
    B, N = 64, 3
    m, q, v, eps, rr, rv = _bucket(B, N)
    cur = torch.cuda.current_device()
    E.analyze_host(m, q, v.copy(), eps, 1.0, "verlet", 10, 0.01, 0, None, None, 0)
    assert torch.cuda.current_device() == cur
    lib = L.load()
    bad = L.HostOpts(flags=L.HOST_ADAPTIVE)                  # adaptive without soft_par
    dyn = np.empty((B, L.N_DYN))
    rc = lib.nb_ensemble_analyze_host_ex(L.ptr(m), L.ptr(q), L.ptr(v), L.ptr(eps), 1.0, B, N, 0, 0, 0.01, 0.01, 0.01, 10, 0,
                                         50, None, None, L.ptr(dyn), None, None, None, 0, 0, ctypes.byref(bad))
    assert rc == -3 and b"soft_par" in lib.nb_last_error()
