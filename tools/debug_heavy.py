import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nbodysimproject_b200 import ensemble as E, _lib as L
from oracle import nbody_oracle as O
rng = np.random.RandomState(17)
N = 4
m = rng.uniform(0.5, 5, N); q = rng.randn(N, 2) * 1.5; q[1] = q[0] + [0.01, 0]; v = rng.randn(N, 2) * 0.3
for mode in ('verlet', 'yoshida4'):
    for use_sort in (False, True):
        o = O.OracleSim(m, q, v, softening=0.02, integrator_mode=mode)
        bk = E.DeviceBucket(m[None], q[None], v[None], 0.02, 1.0, mode)
        bk.prepare(L.PREP_REMOVE_COM | L.PREP_CTOR_KICK, 0.01, 0.01, 0.01)
        if use_sort:
            bk.sort()
        print(mode, 'sort', use_sort, 'n_sub', int(bk.n_sub[0]), o.n_sub_for(0.01), 'v0 err', np.max(np.abs(bk.v.cpu().numpy()[0] - o.v)))
        done = 0
        for target in (1, 2, 5, 20, 50):
            bk.run(0.01, target - done, 0, 0, flags=L.RUN_WRITE_STATE, want_dyn=False)
            for _ in range(target - done):
                o.step(0.01)
            done = target
            print('   steps', target, 'q err', np.max(np.abs(bk.q.cpu().numpy()[0] - o.q)), 'v err', np.max(np.abs(bk.v.cpu().numpy()[0] - o.v)))
