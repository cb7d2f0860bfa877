"""GPU parity: the large-N ham_soft flow (tile-streamed O(N^2) passes, fp32 pair arithmetic) against the dense
fp64 oracle (oracle/largen_hamsoft_oracle.py = the pinned HamSoftOracleSim with the analytic-gradient branch forced).
Tolerances are fp32-level and stated per check."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _system(n, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    m = rng.uniform(0.5, 1.5, n) / n
    q = rng.standard_normal((n, 2)) * scale
    v = rng.standard_normal((n, 2)) * 0.3
    return m, q, v


def _dense(q):
    d = q[:, None, :] - q[None, :, :]
    r2 = d[..., 0] ** 2 + d[..., 1] ** 2
    return d, r2


@pytest.mark.parametrize("n", [2, 3, 777, 1000, 2049])
def test_passes_vs_dense_numpy(n):
    import torch
    from nbodysimproject_b200 import largen as LN
    m, q, v = _system(n, n)
    sim = LN.LargeNSimulation(m, q, v, softening=0.03)
    xym = sim.xym.cpu().numpy().astype(np.float64)
    q32, m32 = xym[:, :2], xym[:, 2]
    d, r2 = _dense(q32)
    off = ~np.eye(n, dtype=bool)
    rng = np.random.default_rng(1)
    h = rng.uniform(0.02, 0.2, n)
    h32 = torch.as_tensor(h.astype(np.float32)).cuda()
    hh = h32.cpu().numpy().astype(np.float64)
    out = torch.zeros((n, 2), dtype=torch.float64, device="cuda")
    jaux = torch.zeros(((n + 1) // 2 * 2, 2), dtype=torch.float32, device="cuda")
    from nbodysimproject_b200 import _lib as L
    lib = L.load()

    def run(kind, iparam=None, eps=0.0):
        L.check(lib.nb_largeN_pass_f32(kind, L.ptr(sim.xym), L.ptr(jaux), n, 0, n, L.ptr(iparam), eps, L.ptr(out),
                                       None, L.stream_ptr()))
        return out.cpu().numpy().copy()

    # DENSITY: S0, S1
    E = np.where(off, np.exp(-r2 / (hh * hh)[:, None]), 0.0)
    S0 = np.sum(m32[None, :] * E, axis=1)
    S1 = np.sum(m32[None, :] * E * r2, axis=1)
    got = run(LN.LN_DENSITY, h32)
    # the self pair (m_i e^0) is summed in fp32 and removed afterwards: absolute floor of a few ulp(m_i)
    assert np.max(np.abs(got[:, 0] - S0)) < 2e-5 * np.max(S0) + 4e-7 * np.max(m32)
    assert np.max(np.abs(got[:, 1] - S1)) < 2e-5 * np.max(S1) + 1e-12
    # EPSGRAD
    A = rng.standard_normal(n)
    nk = -1.4426950408889634 / (hh * hh)
    ja = np.stack([nk, A], 1).astype(np.float32)
    jaux[:n] = torch.as_tensor(ja).cuda()
    nk64, A64 = ja[:, 0].astype(np.float64), ja[:, 1].astype(np.float64)
    Ei = np.where(off, np.exp2(r2 * nk64[:, None]), 0.0)
    Ej = np.where(off, np.exp2(r2 * nk64[None, :]), 0.0)
    Wt = A64[:, None] * m32[None, :] * Ei + A64[None, :] * m32[:, None] * Ej
    gref = np.einsum("ij,ijk->ik", Wt, d)
    got = run(LN.LN_EPSGRAD)
    assert np.max(np.abs(got - gref)) < 5e-5 * np.max(np.abs(gref)) + 1e-12
    # UNITGRAD
    with np.errstate(divide="ignore"):
        w3 = np.where(off & (r2 > 0), r2 ** -1.5, 0.0)
    uref = np.einsum("ij,ijk->ik", w3, d)
    got = run(LN.LN_UNITGRAD)
    assert np.max(np.abs(got - uref)) < 5e-5 * np.max(np.abs(uref)) + 1e-12
    # TAUMIN
    eps = 0.03
    val = np.where(off, (r2 + np.float32(eps) ** 2) ** 1.5 / (m32[:, None] + m32[None, :]), np.inf)
    tref = val.min(axis=1)
    got = run(LN.LN_TAUMIN, eps=eps)[:, :].reshape(-1)[:n]
    assert np.max(np.abs(got - tref) / tref) < 1e-5


@pytest.fixture(scope="module")
def pair():
    """A 300-particle system on both sides; the oracle sees the fp32-rounded inputs the GPU holds."""
    from nbodysimproject_b200.largen import LargeNHamSoftSimulation
    from oracle.largen_hamsoft_oracle import LargeNHamSoftOracle
    m, q, v = _system(300, 11)
    gpu = LargeNHamSoftSimulation(m, q, v, softening=0.02)
    xym = gpu.xym.cpu().numpy().astype(np.float64)
    v32 = gpu.vel.cpu().numpy().astype(np.float64)
    ora = LargeNHamSoftOracle(xym[:, 2], xym[:, :2], v32, softening=0.02, skip_cm_recenter=True)
    return gpu, ora


def test_constructor_calibration(pair):
    """hamsoft_eps_model.py:645-729 + hamiltonian_softening_integrator.py:251-296, 986-1221."""
    gpu, ora = pair
    assert abs(gpu.eps_min - ora.eps_min) < 1e-5 * ora.eps_min
    assert abs(gpu.eps_max - ora.eps_max) < 1e-12
    assert abs(gpu.alpha_run - ora.alpha_run) < 1e-5 * ora.alpha_run
    assert abs(gpu.mu_soft - ora.mu_soft) < 1e-4 * ora.mu_soft
    assert abs(gpu.h_theta - ora.taps["h_theta"]) < 1e-4 * ora.taps["h_theta"]
    assert abs(gpu.h_pi - ora.taps["h_pi"]) < 2e-3 * ora.taps["h_pi"]
    assert abs(gpu.frozen_n_sub - ora.frozen_n_sub) <= 1


def test_eps_star_and_grad(pair):
    gpu, ora = pair
    es_g, g_g = gpu.eps_star_and_grad()
    es_o, g_o = ora.eps_star_and_grad(ora.q)
    assert gpu.last_sweeps == ora.sweeps
    assert abs(es_g - es_o) < 1e-5 * abs(es_o)
    g_g = g_g.cpu().numpy()
    assert np.max(np.abs(g_g - g_o)) < 2e-4 * np.max(np.abs(g_o))


def test_strang_substeps_track_oracle(pair):
    """Three Strang sub-steps S V T V S (hamsoft_stepper.py:247-308) of size h = dt / n_sub."""
    gpu, ora = pair
    h = 0.01 / ora.frozen_n_sub
    for _ in range(3):
        gpu.strang_step(h)
        ora.strang_step(h)
    qg = gpu.xym[:, :2].cpu().numpy().astype(np.float64)
    vg = gpu.vel.cpu().numpy().astype(np.float64)
    assert np.max(np.abs(qg - ora.q)) < 2e-6 * np.max(np.abs(ora.q))
    assert np.max(np.abs(vg - ora.v)) < 2e-5 * np.max(np.abs(ora.v))
    assert abs(gpu.eps - ora.eps) < 1e-5 * abs(ora.eps)
    assert abs(gpu.pi - ora.pi) < 2e-3 * max(abs(ora.pi), 1e-6)


def test_extended_hamiltonian_is_conserved():
    """H_ext = T + U + pi^2/2mu + k/2 (eps-eps*)^2 + S_bar stays put over full macro steps at N = 4096."""
    from nbodysimproject_b200.largen import LargeNHamSoftSimulation, make_disc
    m, q, v = make_disc(4096, seed=2)
    sim = LargeNHamSoftSimulation(m, q, v, softening=0.02, initial_dt=2e-3)
    H0 = sim.extended_hamiltonian()
    P0 = sim.momentum()
    for _ in range(2):
        sim.step(2e-3)
    H1 = sim.extended_hamiltonian()
    P1 = sim.momentum()
    assert abs(H1 - H0) < 2e-4 * abs(H0)
    assert np.all(np.abs(P1[:2] - P0[:2]) < 1e-5)
    assert sim.eps_min <= sim.eps <= sim.eps_max


def test_locality_culling_drops_only_exact_zeros_and_storage_order_is_invisible():
    """The DENSITY / EPSGRAD passes skip (i-block, j-tile) pairs whose bounding boxes are farther apart than 9.35 h
    (ex2.approx.ftz is exactly zero there).  Culled and unculled passes agree to the order of the cross-chunk fp64
    atomics (1e-14); a Morton-sorted simulation gives the caller's particle order back and tracks the unsorted one to
    fp32 summation-order noise."""
    import math
    import torch
    from nbodysimproject_b200 import largen as LN
    n = 20000
    m, q, v = LN.make_disc(n, seed=3)
    soft = 2.0 / math.sqrt(n)
    a = LN.LargeNHamSoftSimulation(m, q, v, softening=soft, initial_dt=1e-3)                         # sorted, culled
    b = LN.LargeNHamSoftSimulation(m, q, v, softening=soft, initial_dt=1e-3, cull=False)             # sorted, every tile
    c = LN.LargeNHamSoftSimulation(m, q, v, softening=soft, initial_dt=1e-3, spatial_sort=False, cull=False)
    assert a.order is not None and c.order is None
    assert np.allclose(a.positions, q, rtol=0, atol=1e-6) and np.allclose(c.positions, q, rtol=0, atol=1e-6)
    # single passes: same h, same jaux -> same sums
    h = torch.full((n,), 0.7 * soft, dtype=torch.float32, device="cuda")
    da = a._pass(LN.LN_DENSITY, h).clone()
    db = b._pass(LN.LN_DENSITY, h).clone()
    assert float((da - db).abs().max()) <= 1e-14 * float(db.abs().max())
    assert float(db.abs().max()) > 0
    es_a, g_a = a.eps_star_and_grad()
    es_b, g_b = b.eps_star_and_grad()
    assert abs(es_a - es_b) <= 1e-13 * abs(es_b)
    assert float((g_a - g_b).abs().max()) <= 1e-12 * float(g_b.abs().max())
    # the culled simulation really skips work: boxes of distant tiles fail the test
    bx = a._tile_boxes(False).cpu().numpy()
    ext = max(bx[:, 2].max() - bx[:, 0].min(), bx[:, 3].max() - bx[:, 1].min())
    assert np.median(np.maximum(bx[:, 2] - bx[:, 0], bx[:, 3] - bx[:, 1])) < 0.35 * ext      # tiles are compact
    hsub = 1e-3 / a.frozen_n_sub
    for sim in (a, b, c):
        for _ in range(2):
            sim.strang_step(hsub)
    assert abs(a.eps - b.eps) <= 1e-12 * abs(b.eps) and abs(a.pi - b.pi) <= 1e-9 * max(abs(b.pi), 1e-9)
    assert np.max(np.abs(a.positions - b.positions)) <= 1e-12
    assert abs(a.eps - c.eps) <= 2e-5 * abs(c.eps)
    assert np.max(np.abs(a.positions - c.positions)) <= 2e-6 * np.max(np.abs(c.positions))
    assert np.max(np.abs(a.velocities - c.velocities)) <= 2e-4 * np.max(np.abs(c.velocities))
    # one long-range UNITGRAD pass per sub-step (shared by the closing and the next opening S half-flow)
    assert a._unit_cache is not None


def test_reflection_policy_folds_epsilon_like_the_oracle():
    """Barrier policy 'reflection' (use_soft_barrier = False): epsilon is folded into [eps_min, eps_max] and pi flipped
    (hamsoft_barrier_controller.py:27-69, hamsoft_utils.py:159-184) at the four places the stepper does it.  A narrow
    admissible interval makes epsilon hit both walls within a few sub-steps."""
    from nbodysimproject_b200.largen import LargeNHamSoftSimulation
    from oracle.largen_hamsoft_oracle import LargeNHamSoftOracle
    m, q, v = _system(200, 5)
    gpu = LargeNHamSoftSimulation(m, q, v, softening=0.02, use_soft_barrier=False, spatial_sort=False)
    xym = gpu.xym.cpu().numpy().astype(np.float64)
    v32 = gpu.vel.cpu().numpy().astype(np.float64)
    ora = LargeNHamSoftOracle(xym[:, 2], xym[:, :2], v32, softening=0.02, skip_cm_recenter=True, use_soft_barrier=False)
    assert gpu.reflect_policy and ora.reflect_policy and not gpu.soft_policy
    for sim in (gpu, ora):                       # squeeze the interval around the start value
        sim.eps_max = float(sim.eps) * 1.002
        sim.eps_min = float(sim.eps) * 0.9995
    h = 0.01 / ora.frozen_n_sub
    flips = 0
    for _ in range(6):
        p_before = ora.pi
        gpu.strang_step(h)
        ora.strang_step(h)
        assert gpu.eps_min <= gpu.eps <= gpu.eps_max
        assert abs(gpu.eps - ora.eps) < 2e-5 * abs(ora.eps)
        assert abs(gpu.pi - ora.pi) < 5e-3 * max(abs(ora.pi), 1e-6)
        flips += int(np.sign(ora.pi) != np.sign(p_before))
    qg = gpu.xym[:, :2].cpu().numpy().astype(np.float64)
    assert np.max(np.abs(qg - ora.q)) < 5e-6 * np.max(np.abs(ora.q))
    assert flips >= 1                            # the walls were really hit
