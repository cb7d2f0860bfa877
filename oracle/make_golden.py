"""Generate golden vectors from the LIVE reference (run in the build container only).

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference (/root/reference, pure Python) is imported with the one-line lightgbm stub from
SURVEY.md section 8c.  It cannot travel to the GPU box, so its outputs on seeded inputs are committed
as small fixtures; tests compare both the oracle (tests/test_oracle_golden.py, CPU) and the CUDA
path (tests/test_gpu_*.py) against them.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

REF = os.environ.get("NBODY_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    sys.dont_write_bytecode = True
    sys.modules.setdefault("lightgbm", types.ModuleType("lightgbm"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import minbody  # noqa: F401
    return minbody


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


# ---------------------------------------------------------------------------------------------
# deterministic test systems (shared with the tests through the .npz files themselves)
# ---------------------------------------------------------------------------------------------

def named_systems(mb):
    rng = np.random.RandomState(1234)
    S = {}
    S["readme3"] = (np.array([1.0, 0.5, 0.1]), np.array([[0, 0], [1, 0], [2, 0.0]]),
                    np.array([[0, 0], [0, 1], [0, 0.5]]), 1e-3)
    m, p, v = mb.SpecializedGenerators.generate_hierarchical_triple(1.0, 0.5, 10.0)
    S["hier3"] = (m, p, v, 0.01)
    m, p, v = mb.SpecializedGenerators.generate_equal_mass_polygon(5, 1.0, 0.5)
    S["poly5"] = (m, p, v, 0.05)
    m, p, v = mb.SpecializedGenerators.generate_equal_mass_polygon(7, 2.0, 0.8)
    S["poly7"] = (m, p, v, 0.05)
    for n in (4, 6, 8):
        m = rng.uniform(0.1, 10.0, n)
        p = rng.randn(n, 2) * 2.0
        v = rng.randn(n, 2) * 0.7
        S[f"rand{n}"] = (m, p, v, 0.05)
    # a close pair -> n_sub > 1
    m = np.array([2.0, 1.0, 0.5, 0.2])
    p = np.array([[0.0, 0.0], [0.12, 0.0], [1.5, 0.3], [-1.0, 1.2]])
    v = np.array([[0.0, -1.2], [0.0, 2.4], [0.3, 0.6], [-0.2, -0.4]])
    S["close4"] = (m, p, v, 0.01)
    return S


def gen_pair(mb):
    from minbody.forces import gravitational_force, dV_d_epsilon
    from minbody.potential import softened_potential
    from minbody.tangent_map import TangentMap
    rng = np.random.RandomState(7)
    out = {}
    case = 0
    for n in (2, 3, 4, 5, 6, 7, 8):
        for eps, G in ((0.0, 1.0), (1e-3, 1.0), (0.05, 1.0), (0.3, 2.5)):
            q = rng.randn(n, 2) * rng.uniform(0.2, 3.0)
            m = rng.uniform(0.1, 10.0, n)
            dr = rng.randn(n, 2)
            F = gravitational_force(q, m, eps=eps, G=G)
            acc = F / m[:, None]
            dV = dV_d_epsilon(q, m, eps, G)
            U = softened_potential(q, m, G, eps)
            with quiet():
                sim = mb.NBodySimulation(masses=m, positions=q, velocities=np.zeros_like(q), G=G,
                                         softening=max(eps, 0.0), integrator_mode="verlet",
                                         skip_init_corrector=True)
            sim.manager._step_s2 = eps * eps
            da = TangentMap(sim).variational_accel(dr)
            k = f"c{case:02d}_"
            out.update({k + "q": q, k + "m": m, k + "eps": eps, k + "G": G, k + "dr": dr, k + "F": F,
                        k + "acc": acc, k + "dV": dV, k + "U": U, k + "da": da})
            case += 1
    out["n_cases"] = case
    np.savez_compressed(os.path.join(OUT, "pair_kernels.npz"), **out)
    print("pair_kernels:", case, "cases")


def gen_kepler(mb):
    from minbody.kepler_solver import UniversalVariableKeplerSolver
    s = UniversalVariableKeplerSolver()
    rng = np.random.RandomState(11)
    R, V, MU, DT, RO, VO = [], [], [], [], [], []
    for i in range(200):
        r = rng.randn(2) * rng.uniform(0.3, 3.0)
        mu = rng.uniform(0.5, 3.0)
        vc = np.sqrt(mu / np.linalg.norm(r))
        kind = i % 4
        if kind == 0:      # near-circular
            t = np.array([-r[1], r[0]]) / np.linalg.norm(r)
            v = t * vc * rng.uniform(0.9, 1.1)
        elif kind == 1:    # eccentric bound
            v = rng.randn(2) * 0.5 * vc
        elif kind == 2:    # hyperbolic
            v = rng.randn(2) * 2.0 * vc
        else:              # tiny planet-like step
            t = np.array([-r[1], r[0]]) / np.linalg.norm(r)
            v = t * vc
        dt = [0.005, 0.0314, 0.1, 0.5][i % 4] * (1 if i % 7 else -1)
        ro, vo = s.propagate(r, v, mu, dt)
        R.append(r); V.append(v); MU.append(mu); DT.append(dt); RO.append(ro); VO.append(vo)
    C = []
    for z in np.concatenate([np.linspace(-30, 30, 61), rng.randn(40) * 0.05]):
        C.append((z,) + tuple(s._cfunc(z)))
    np.savez_compressed(os.path.join(OUT, "kepler.npz"), r=np.array(R), v=np.array(V), mu=np.array(MU),
                        dt=np.array(DT), r_out=np.array(RO), v_out=np.array(VO), cfunc=np.array(C))
    print("kepler: 200 cases")


def _alt_force(q, m, eps=0.0, G=1.0):
    """Mathematically identical to forces.py:63-75 with different rounding (sequential pair loop,
    (1/sqrt)^3, no m_i multiply/divide) -- what any re-implementation, CPU or GPU, looks like."""
    q = np.asarray(q, dtype=float)
    m = np.asarray(m, dtype=float)
    n = len(m)
    F = np.zeros_like(q)
    if n < 2 or G == 0.0:
        return F
    for i in range(n):
        for j in range(i + 1, n):
            d = q[i] - q[j]
            r2 = d[0] * d[0] + d[1] * d[1] + eps * eps
            if r2 > 0.0:
                w = 1.0 / np.sqrt(r2)
                f = G * m[i] * m[j] * w * w * w * d
                F[i] -= f
                F[j] += f
    return F


def _sensitivity(mb, m, p, v, soft, mode, dt, targets):
    """Divergence of the REFERENCE from itself when its force routine is replaced by _alt_force:
    the chaotic amplification of last-bit rounding, i.e. the horizon beyond which no tolerance is meaningful."""
    import minbody.simulation as simmod
    with quiet():
        a = mb.NBodySimulation(masses=m, positions=p, velocities=v, softening=soft, integrator_mode=mode)
    orig = simmod.gravitational_force
    simmod.gravitational_force = _alt_force
    try:
        with quiet():
            b = mb.NBodySimulation(masses=m, positions=p, velocities=v, softening=soft, integrator_mode=mode)
        out = {}
        done = 0
        for t in targets:
            for _ in range(t - done):
                simmod.gravitational_force = orig
                a.step(dt)
                simmod.gravitational_force = _alt_force
                b.step(dt)
            done = t
            out[t] = float(np.max(np.abs(a._pos - b._pos)) / np.max(np.abs(a._pos)))
    finally:
        simmod.gravitational_force = orig
    return out


def gen_traj(mb):
    S = named_systems(mb)
    out = {}
    names = []
    for name, (m, p, v, soft) in S.items():
        for mode in ("verlet", "yoshida4"):
            with quiet():
                sim = mb.NBodySimulation(masses=m, positions=p, velocities=v, softening=soft,
                                         integrator_mode=mode)
            key = f"{name}_{mode}_"
            names.append(key)
            out[key + "m"] = m
            out[key + "q_in"] = p
            out[key + "v_in"] = v
            out[key + "soft"] = soft
            out[key + "v0"] = sim._vel.copy()            # after COM removal + ctor half kick
            out[key + "h_sub_ref"] = sim._integrator.h_sub_ref
            dt = 0.01
            done = 0
            for target in (1, 10, 100, 1000):
                for _ in range(target - done):
                    sim.step(dt)
                done = target
                out[key + f"q{target}"] = sim._pos.copy()
                out[key + f"v{target}"] = sim._vel.copy()
            # Integrator.step counts each sub-step once (integrator.py:98-100); verlet's atomicstep
            # counts it a second time (integrator.py:113)
            cnt = sim._integrator._substeps_in_last_step
            out[key + "n_sub"] = cnt // 2 if mode == "verlet" else cnt
            with quiet():
                snap = sim.snapshot()
            out[key + "v_snap"] = sim._vel.copy()         # snapshot half kick applied
            for t, val in _sensitivity(mb, m, p, v, soft, mode, dt, (1, 10, 100, 1000)).items():
                out[key + f"sens{t}"] = val
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "trajectories.npz"), **out)
    print("trajectories:", len(names))


def whfast_systems():
    rng = np.random.RandomState(99)
    S = {}
    for k, npl in enumerate((2, 3, 4)):
        m = np.concatenate([[1.0], 10 ** rng.uniform(-6, -3, npl)])
        a = np.cumsum(np.concatenate([[1.0], rng.uniform(0.3, 0.8, npl - 1)]))
        ph = rng.uniform(0, 2 * np.pi, npl)
        p = np.zeros((npl + 1, 2))
        v = np.zeros((npl + 1, 2))
        for i in range(npl):
            p[i + 1] = a[i] * np.array([np.cos(ph[i]), np.sin(ph[i])])
            vc = np.sqrt((m[0] + m[i + 1]) / a[i])
            v[i + 1] = vc * np.array([-np.sin(ph[i]), np.cos(ph[i])])
        S[f"planets{npl}"] = (m, p, v)
    # heavier companions (dominance still > 0.2), eccentric
    m = np.array([1.0, 0.05, 0.01])
    p = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 2.2]])
    v = np.array([[0.0, 0.0], [0.0, 1.15], [-0.6, 0.0]])
    S["heavy3"] = (m, p, v)
    return S


def gen_whfast(mb):
    out = {}
    names = []
    for name, (m, p, v) in whfast_systems().items():
        with quiet():
            sim = mb.NBodySimulation(masses=m, positions=p, velocities=v, softening=0.0,
                                     integrator_mode="whfast")
        assert sim._integrator_mode == "whfast"
        key = name + "_"
        names.append(key)
        out[key + "m"] = m; out[key + "q_in"] = p; out[key + "v_in"] = v
        out[key + "v0"] = sim._vel.copy()
        out[key + "h_sub_ref"] = sim._integrator.h_sub_ref
        dt = 0.01 * 2 * np.pi
        done = 0
        for target in (1, 10, 100, 500):
            for _ in range(target - done):
                sim.step(dt)
            done = target
            out[key + f"q{target}"] = sim._pos.copy()
            out[key + f"v{target}"] = sim._vel.copy()
        out[key + "n_sub"] = sim._integrator._substeps_in_last_step
    out["names"] = np.array(names)
    out["dt"] = 0.01 * 2 * np.pi
    np.savez_compressed(os.path.join(OUT, "whfast.npz"), **out)
    print("whfast:", len(names))


def gen_features(mb):
    """BatchStabilityAnalyzer 'full' rows for a seeded batch, with the tangent draws recorded."""
    S = named_systems(mb)
    for mode, n_steps in (("verlet", 300), ("yoshida4", 1000)):
        sims = []
        keys = []
        for name, (m, p, v, soft) in S.items():
            with quiet():
                sims.append(mb.NBodySimulation(masses=m, positions=p, velocities=v, softening=soft,
                                               integrator_mode=mode))
            keys.append(name)
        draws = []
        orig = np.random.randn

        def rec(*shape):
            a = orig(*shape)
            draws.append(a.copy())
            return a

        np.random.seed(2024)
        np.random.randn = rec
        try:
            with quiet():
                df = mb.BatchStabilityAnalyzer(n_steps=n_steps, dt=0.01, mode="full").analyze_batch(sims, show_progress=False)
        finally:
            np.random.randn = orig
        # sensitivity: the same batch, same tangent draws, with the reference's force routine swapped for the
        # equivalent-arithmetic one
        import minbody.simulation as simmod
        it = iter([d.copy() for d in draws])
        sims2 = []
        orig_force = simmod.gravitational_force
        simmod.gravitational_force = _alt_force
        np.random.randn = lambda *shape: next(it)
        try:
            for name in keys:
                m, p, v, soft = S[name]
                with quiet():
                    sims2.append(mb.NBodySimulation(masses=m, positions=p, velocities=v, softening=soft,
                                                    integrator_mode=mode))
            with quiet():
                df2 = mb.BatchStabilityAnalyzer(n_steps=n_steps, dt=0.01, mode="full").analyze_batch(sims2, show_progress=False)
        finally:
            np.random.randn = orig
            simmod.gravitational_force = orig_force
        out = {"names": np.array(keys), "n_steps": n_steps, "dt": 0.01, "columns": np.array(list(df.columns))}
        for i, name in enumerate(keys):
            for c in df.columns:
                a, b = df.iloc[i][c], df2.iloc[i][c]
                if isinstance(a, (str, bool, np.bool_)):
                    continue
                d = abs(float(a) - float(b))
                out[f"{name}__sens__{c}"] = d if np.isfinite(d) else 0.0
        for i, name in enumerate(keys):
            m, p, v, soft = S[name]
            out[f"{name}_m"] = m; out[f"{name}_q"] = p; out[f"{name}_v"] = v; out[f"{name}_soft"] = soft
            out[f"{name}_raw_r"] = draws[2 * i]
            out[f"{name}_raw_v"] = draws[2 * i + 1]
            row = df.iloc[i]
            for c in df.columns:
                val = row[c]
                if isinstance(val, (str, bool, np.bool_)):
                    out[f"{name}__{c}"] = np.array(str(val))
                else:
                    out[f"{name}__{c}"] = float(val)
        np.savez_compressed(os.path.join(OUT, f"features_{mode}.npz"), **out)
        print("features", mode, df.shape)


def main():
    os.makedirs(OUT, exist_ok=True)
    mb = import_reference()
    which = sys.argv[1:] or ["pair", "kepler", "traj", "whfast", "features", "hamsoft"]
    if "pair" in which:
        gen_pair(mb)
    if "kepler" in which:
        gen_kepler(mb)
    if "traj" in which:
        gen_traj(mb)
    if "whfast" in which:
        gen_whfast(mb)
    if "features" in which:
        gen_features(mb)
    if "hamsoft" in which:
        try:
            from make_golden_hamsoft import gen_hamsoft
        except ImportError:
            sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
            try:
                from make_golden_hamsoft import gen_hamsoft
            except ImportError:
                gen_hamsoft = None
        if gen_hamsoft is not None:
            gen_hamsoft(mb, OUT)


if __name__ == "__main__":
    main()
