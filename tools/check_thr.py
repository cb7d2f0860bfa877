import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodysimproject_b200 import ensemble as E, _lib as L
from nbodysimproject_b200.generators import EnsembleInputs
rng = np.random.default_rng(42)
buckets = EnsembleInputs.diverse(rng, 1 << 18, n_max=8)
for N, (m, q, v, soft, cohort) in sorted(buckets.items()):
    bk = E.DeviceBucket(m, q, v, soft, 1.0, "yoshida4")
    bk.prepare(L.PREP_REMOVE_COM | L.PREP_CTOR_KICK | L.PREP_SNAPSHOT_KICK, 0.01, 0.01, 0.01, 50)
    bk.sort()
    ns = bk.n_sub.cpu().numpy()
    print("N", N, "B", bk.B, "W", int(ns.sum()), "n_heavy", int(bk._bins[64]), "thr", int(bk._bins[65]),
          "hist>4:", int((ns > 4).sum()), ">10:", int((ns > 10).sum()), ">=50:", int((ns >= 50).sum()),
          flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0, v0 = bk.q.clone(), bk.v.clone()
    for rep in range(2):
        bk.q.copy_(q0); bk.v.copy_(v0)
        e0.record(); bk.run(0.01, 1000, 10, 0, flags=0, want_dyn=False); e1.record(); torch.cuda.synchronize()
    print("   main kernel solo ms", round(e0.elapsed_time(e1), 2))
