import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nbodysimproject_b200 import ensemble as E, _lib as L
from oracle import nbody_oracle as O
g = np.load('tests/golden/features_verlet.npz')
name = 'rand6'
m, q, v, soft = g[f'{name}_m'], g[f'{name}_q'], g[f'{name}_v'], float(g[f'{name}_soft'])
sim = O.OracleSim(m, q, v, softening=soft, integrator_mode='verlet')
c = sim.snapshot_restore()
bk = E.DeviceBucket(m[None], q[None], v[None], soft, 1.0, 'verlet')
bk.prepare(L.PREP_REMOVE_COM | L.PREP_CTOR_KICK, 0.01, 0.01, 0.01)
bk.prepare(L.PREP_SNAPSHOT_KICK, 0.01, 0.01, 0.01)
print('v after kicks relerr', np.max(np.abs(bk.v.cpu().numpy()[0] - c.v)) / np.max(np.abs(c.v)), 'n_sub', int(bk.n_sub[0]), c.n_sub_for(0.01))
for i in range(300):
    c.step(0.01)
    bk.run(0.01, 1, 0, 0, flags=L.RUN_WRITE_STATE, want_dyn=False)
    if i % 30 == 0 or i == 299:
        qq, vv = bk.q.cpu().numpy()[0], bk.v.cpu().numpy()[0]
        Lg = m * (qq[:, 0] * vv[:, 1] - qq[:, 1] * vv[:, 0]); Lo = c.m * (c.q[:, 0] * c.v[:, 1] - c.q[:, 1] * c.v[:, 0])
        print(i, 'q', np.max(np.abs(qq - c.q)) / np.max(np.abs(c.q)), 'v', np.max(np.abs(vv - c.v)) / np.max(np.abs(c.v)), 'var', np.var(Lg), np.var(Lo))
# one-shot run with samples
bk2 = E.DeviceBucket(m[None], q[None], v[None], soft, 1.0, 'verlet')
bk2.prepare(L.PREP_REMOVE_COM | L.PREP_CTOR_KICK, 0.01, 0.01, 0.01)
bk2.prepare(L.PREP_SNAPSHOT_KICK, 0.01, 0.01, 0.01)
for interval in (1, 3):
    bk3 = E.DeviceBucket(m[None], q[None], bk2.v, soft, 1.0, 'verlet')
    dyn = bk3.run(0.01, 300, interval, 0, flags=L.RUN_ENERGY).cpu().numpy()[0]
    print('interval', interval, dict(zip(L.DYN_COLUMNS, dyn))['ang_mom_var_mean'])
