"""Short driver for ncu: one DENSITY and one EPSGRAD pass of the large-N ham_soft flow.  python tools/profile_pass.py [n]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodysimproject_b200 import largen as LN

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
m, q, v = LN.make_disc(n, 1)
sim = LN.LargeNHamSoftSimulation(m, q, v, softening=2.0 / np.sqrt(n), initial_dt=1e-3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
h = torch.full((n,), float(sim.eps), dtype=torch.float32, device="cuda")
for kind, name in ((LN.LN_DENSITY, "DENSITY"), (LN.LN_EPSGRAD, "EPSGRAD"), (LN.LN_UNITGRAD, "UNITGRAD")):
    for rep in range(3):
        e0.record(); sim._pass(kind, h if kind == LN.LN_DENSITY else None); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    print(f"pass {name} n={n}: {t*1e3:.2f} ms, {float(n)*n/t:.3e} pair evaluations/s")
