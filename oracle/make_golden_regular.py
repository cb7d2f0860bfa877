"""Golden vectors on REGULAR systems with N = 5..8 bodies, from the live reference (the chaotic
rand6 / rand8 goldens need tolerance windows of +-4 on some feature columns, which tests nothing).  A heavy central body
with light companions on near-circular, well separated, softened orbits does not amplify rounding noise (the recorded
self-sensitivities are ~1e-13 after 1000 steps), so every trajectory entry and every feature column is compared at the
fixed floors of the tests, for every N.  Same generators as oracle/make_golden.py, different system set.
Run in the build container:  python oracle/make_golden_regular.py
  ->  tests/golden/trajectories_regular.npz, features_regular_verlet.npz, features_regular_yoshida4.npz"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as MG  # noqa: E402

OUT = MG.OUT


def regular_systems(mb):
    S = {}
    rng = np.random.RandomState(77)
    for n in (5, 6, 7, 8):
        m = np.concatenate([[1.0], rng.uniform(0.5, 2.0, n - 1) * 1e-3])
        r = 1.0 + 0.45 * np.arange(n - 1) + rng.uniform(-0.03, 0.03, n - 1)
        ph = rng.uniform(0, 2 * np.pi, n - 1)
        p = np.concatenate([[[0.0, 0.0]], np.stack([r * np.cos(ph), r * np.sin(ph)], 1)])
        vc = np.sqrt(1.0 / r) * (1.0 + rng.uniform(-0.01, 0.01, n - 1))
        v = np.concatenate([[[0.0, 0.0]], np.stack([-vc * np.sin(ph), vc * np.cos(ph)], 1)])
        S[f"ring{n}"] = (m, p, v, 0.02)
    # a regular system that needs sub-steps: two light bodies 0.004 apart (the unsoftened schedule asks for n_sub = 3)
    # whose mutual force the softening (0.2) keeps smooth
    m, p, v, _ = [a.copy() if hasattr(a, "copy") else a for a in S["ring6"]]
    p[5] = p[4] + np.array([0.004, 0.0])
    v[5] = v[4]
    S["ring6sub"] = (m, p, v, 0.2)
    return S


def main():
    mb = MG.import_reference()
    tmp = tempfile.mkdtemp()
    MG.named_systems = regular_systems
    MG.OUT = tmp
    try:
        MG.gen_traj(mb)
        MG.gen_features(mb)
        shutil.move(os.path.join(tmp, "trajectories.npz"), os.path.join(OUT, "trajectories_regular.npz"))
        for mode in ("verlet", "yoshida4"):
            shutil.move(os.path.join(tmp, f"features_{mode}.npz"), os.path.join(OUT, f"features_regular_{mode}.npz"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    g = np.load(os.path.join(OUT, "trajectories_regular.npz"))
    for key in g["names"]:
        key = str(key)
        print(key, "n_sub", int(g[key + "n_sub"]), "self-sensitivity @1000:", float(g[key + "sens1000"]))
    for mode in ("verlet", "yoshida4"):
        f = np.load(os.path.join(OUT, f"features_regular_{mode}.npz"))
        worst = max(float(f[k]) for k in f.files if "__sens__" in k)
        print("features", mode, "largest self-sensitivity of any column:", worst)


if __name__ == "__main__":
    main()
