"""GPU, BASELINE.json's full sizes: properties that do not need the O(seconds-per-system) oracle.

  * C3 ensemble (2^20 systems, N = 3..8, yoshida4): linear momentum conserved to machine precision, angular
    momentum to rounding, time reversibility of the symmetric KDK schemes (forward n steps, backward n steps
    returns to the start), identical results for a re-ordered batch, zero non-finite statuses;
  * C5 large N (N = 2^20): total force vanishes to fp32 accuracy, accelerations of a random subset of particles
    equal an fp64 direct sum over all 2^20 sources (oracle formula, forces.py:63-75), and the packed-f32x2 kernel
    agrees with the scalar fp32 kernel.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

B_FULL = 1 << 20


@pytest.fixture(scope="module")
def cohort():
    from nbodysimproject_b200.generators import EnsembleInputs
    rng = np.random.default_rng(2026)
    return EnsembleInputs.diverse(rng, B_FULL, n_max=8)


def test_c3_full_size_conservation_and_reversibility(cohort):
    import torch
    from nbodysimproject_b200 import _lib as L
    from nbodysimproject_b200 import ensemble as E
    total = 0
    for N, (m, q, v, soft, _) in sorted(cohort.items()):
        bk = E.DeviceBucket(m, q, v, soft, 1.0, "yoshida4")
        bk.prepare(L.PREP_REMOVE_COM, 0.01, 0.01, 0.01, 50)       # COM frame, no corrector kick: a pure KDK map follows
        bk.sort()
        q0, v0 = bk.q.clone(), bk.v.clone()
        md = bk.m
        P0 = (md[:, :, None] * v0).sum(1)
        L0 = (md * (q0[:, :, 0] * v0[:, :, 1] - q0[:, :, 1] * v0[:, :, 0])).sum(1)
        bk.run(0.01, 100, 0, 0, flags=L.RUN_WRITE_STATE, want_dyn=False)
        assert int((bk.status != 0).sum()) == 0
        P1 = (md[:, :, None] * bk.v).sum(1)
        L1 = (md * (bk.q[:, :, 0] * bk.v[:, :, 1] - bk.q[:, :, 1] * bk.v[:, :, 0])).sum(1)
        pscale = (md[:, :, None] * bk.v.abs()).sum(1).clamp_min(1e-300)
        assert float(((P1 - P0).abs() / pscale).max()) < 1e-10          # 100 steps x <= 50 sub-steps of rounding
        lscale = (md * (bk.q[:, :, 0] * bk.v[:, :, 1]).abs() + md * (bk.q[:, :, 1] * bk.v[:, :, 0]).abs()).sum(1).clamp_min(1e-300)
        assert float(((L1 - L0).abs() / lscale).max()) < 1e-9
        # time reversibility: the same map with dt -> -dt undoes the 100 steps
        bk.run(-0.01, 100, 0, 0, flags=L.RUN_WRITE_STATE, want_dyn=False)
        scale = q0.abs().amax(dim=(1, 2)).clamp_min(1e-300)
        err = ((bk.q - q0).abs().amax(dim=(1, 2)) / scale)
        # chaotic close encounters amplify rounding (and a reversed step re-rounds every kick): the bulk must return
        # to the start to ~1e-11, and only a small tail of strongly interacting systems may drift further
        assert bool(torch.isfinite(err).all())
        assert float(err.median()) < 1e-11, float(err.median())
        frac8 = float((err < 1e-8).double().mean())
        frac4 = float((err < 1e-4).double().mean())
        assert frac8 > 0.80 and frac4 > 0.95, (N, frac8, frac4, float(err.max()))   # N = 8: 0.865 / 0.975
        total += bk.B
    assert total == B_FULL


def test_c3_full_size_order_invariance(cohort):
    """Per-system results do not depend on where a system sits in the batch (sorting, warp packing, heavy mappings)."""
    import torch
    from nbodysimproject_b200 import _lib as L
    from nbodysimproject_b200 import ensemble as E
    N = 6
    m, q, v, soft, _ = cohort[N]
    rng = np.random.default_rng(1)
    p = rng.permutation(m.shape[0])
    outs = []
    for mm, qq, vv, ss in ((m, q, v, soft), (m[p], q[p], v[p], soft[p])):
        rr = np.random.default_rng(3).standard_normal((m.shape[0], N, 2))
        rv = np.random.default_rng(4).standard_normal((m.shape[0], N, 2))
        if mm is not m:
            rr, rv = rr[p], rv[p]
        bk = E.DeviceBucket(mm, qq, vv, ss, 1.0, "yoshida4")
        bk.prepare(L.PREP_REMOVE_COM | L.PREP_CTOR_KICK | L.PREP_SNAPSHOT_KICK, 0.01, 0.01, 0.01, 50, want_static=True)
        bk.sort()
        dyn = bk.run(0.01, 200, 2, 20, rr, rv, flags=L.RUN_ENERGY)
        outs.append((dyn.cpu().numpy(), bk.static.cpu().numpy()))
    assert np.array_equal(outs[0][0][p], outs[1][0], equal_nan=True)
    assert np.array_equal(outs[0][1][p], outs[1][1], equal_nan=True)


def test_c5_full_size_force_properties():
    import torch
    from nbodysimproject_b200 import _lib as L
    from nbodysimproject_b200.largen import LargeNSimulation, make_disc
    n = 1 << 20
    m, q, v = make_disc(n, seed=1)
    sim = LargeNSimulation(m, q, v, softening=1e-3)
    acc = sim.accelerations().double()
    xym = sim.xym.double()
    F = (xym[:, 2:3] * acc).sum(0)
    scale = (xym[:, 2:3] * acc.abs()).sum()
    assert float(F.abs().max() / scale) < 1e-6                      # Newton's third law, fp32 pair arithmetic
    # fp64 direct sum for 48 random particles against all 2^20 sources (forces.py:63-75)
    idx = torch.as_tensor(np.random.default_rng(0).choice(n, 48, replace=False)).cuda()
    d = xym[idx, None, :2] - xym[None, :, :2]
    r2 = (d * d).sum(-1) + 1e-3 ** 2
    w3 = r2 ** -1.5
    w3[torch.arange(48, device="cuda"), idx] = 0.0
    ref = -(xym[None, :, 2:3] * w3[:, :, None] * d).sum(1)
    rms = ref.pow(2).mean().sqrt()
    assert float((acc[idx] - ref).abs().max() / rms) < 2e-5
    # scalar fp32 kernel (v1) and packed f32x2 kernel (default) agree
    sim.variant = 1
    acc1 = sim.accelerations().double().clone()
    sim.variant = -1
    assert float((acc1 - acc).abs().max() / acc.abs().max()) < 1e-5
