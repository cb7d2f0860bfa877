"""GPU: on-device initial conditions.  The RNG differs from NumPy's, so parity with the host
generators (which mirror the reference's distributions) is statistical: moments and invariants, plus exact structural
properties and index-only reproducibility (sharding invariance)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _np(*ts):
    return [t.cpu().numpy() for t in ts]


def test_index_only_reproducibility():
    from nbodysimproject_b200.generators import generate_on_device
    full = _np(*generate_on_device("random", 5, 3000, seed=7))
    a = _np(*generate_on_device("random", 5, 1000, seed=7, first_index=0))
    b = _np(*generate_on_device("random", 5, 2000, seed=7, first_index=1000))
    for f, x, y in zip(full, a, b):
        assert np.array_equal(f, np.concatenate([x, y]))
    other = _np(*generate_on_device("random", 5, 3000, seed=8))
    assert not np.array_equal(full[1], other[1])


def test_random_cohort_matches_host_distributions():
    from nbodysimproject_b200.generators import EnsembleInputs, generate_on_device
    B, N = 200000, 4
    m, q, v, eps = _np(*generate_on_device("random", N, B, seed=3))
    hm, hq, hv, heps = EnsembleInputs.random(np.random.default_rng(3), B, N)
    assert np.all((m >= 0.1) & (m <= 10.0)) and np.all((eps >= 0.001) & (eps <= 0.1))
    # uniform / log-uniform mass halves, position scale, softening: first two moments within 1 %
    for g, h in ((m[1::2], hm[1::2]), (np.log(m[0::2]), np.log(hm[0::2])), (eps, heps)):
        assert abs(g.mean() - h.mean()) < 0.01 * abs(h.mean()) + 0.01 * h.std()
        assert abs(g.std() - h.std()) < 0.02 * h.std()
    assert abs(q.std() - hq.std()) < 0.01 * hq.std() and abs(q.mean()) < 0.01
    # centre-of-mass velocity removed exactly; virial ratio distribution matches the host generator
    P = (m[:, :, None] * v).sum(1)
    assert np.max(np.abs(P)) < 1e-12 * np.max(np.abs(m[:, :, None] * v))

    def virial(m, q, v, eps):
        d = q[:, :, None, :] - q[:, None, :, :]
        r = np.sqrt((d ** 2).sum(-1))
        iu = np.triu_indices(N, 1)
        U = -np.sum(m[:, iu[0]] * m[:, iu[1]] / (r[:, iu[0], iu[1]] + eps[:, None]), axis=1)
        T = 0.5 * (m * (v ** 2).sum(-1)).sum(1)
        return 2 * T / np.abs(U)

    vg, vh = virial(m, q, v, eps), virial(hm, hq, hv, heps)
    assert abs(np.median(vg) - np.median(vh)) < 0.02 * np.median(vh)
    assert abs(np.percentile(vg, 90) - np.percentile(vh, 90)) < 0.05 * np.percentile(vh, 90)


def test_structured_cohorts():
    from nbodysimproject_b200.generators import generate_on_device
    m, q, v, eps = _np(*generate_on_device("polygon", 6, 1000, seed=1))
    r = np.sqrt((q ** 2).sum(-1))
    assert np.allclose(r, r[:, :1], rtol=1e-14) and np.all(m == 1.0) and np.all(eps == 0.05)
    assert np.all((r[:, 0] >= 0.5) & (r[:, 0] <= 3.0))
    assert np.allclose((q * v).sum(-1), 0.0, atol=1e-13)                       # tangential velocities
    m, q, v, eps = _np(*generate_on_device("hierarchical", 3, 1000, seed=1))
    assert np.all(m[:, 0] == 1.0) and np.all((m[:, 1] >= 0.1) & (m[:, 1] <= 1.0)) and np.all(q[:, 2, 0] >= 5.0)
    assert np.allclose(q[:, 1, 0] - q[:, 0, 0], 1.0, rtol=1e-14) and np.all(eps == 0.01)
    m, q, v, eps = _np(*generate_on_device("planetary_ttv", 4, 1000, seed=1))
    a = np.sqrt((q[:, 1:] ** 2).sum(-1))
    sp = np.sqrt((v[:, 1:] ** 2).sum(-1))
    assert np.allclose(a[:, 0], 1.0, rtol=1e-14) and np.allclose(sp, np.sqrt((1 + m[:, 1:]) / a), rtol=1e-13)
    ratio = (a[:, 1] / a[:, 0]) ** 1.5
    near = np.min(np.abs(ratio[:, None] / np.array([1.5, 2.0, 5 / 3])[None, :] - 1.0), axis=1)
    assert np.all(near <= 0.02 + 1e-12) and np.all(eps == 0.0)
    assert np.all(m[:, 1] >= 1e-5 * (1 - 1e-12)) and np.all(m[:, 2:] <= 1e-3 * (1 + 1e-12))


def test_generated_systems_run_through_the_ensemble_kernels():
    from nbodysimproject_b200 import _lib as L, ensemble as E
    from nbodysimproject_b200.generators import generate_on_device
    m, q, v, eps = generate_on_device("close", 3, 4096, seed=9)
    bk = E.DeviceBucket(m, q, v, eps, 1.0, "yoshida4")
    bk.prepare(L.PREP_REMOVE_COM | L.PREP_CTOR_KICK, 0.01, 0.01, 0.01, 50)
    bk.sort()
    bk.run(0.01, 50, 0, 0, flags=L.RUN_WRITE_STATE, want_dyn=False)
    assert int((bk.status != 0).sum()) == 0
