"""Short driver for `ncu --set full`: the dominant kernel (ensemble_main_kernel<N, yoshida4>, thread-per-system)
on a uniform batch (n_sub = 1, so no sub-step tail), plus the large-N kernel.  python tools/profile_main.py [N] [B] [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodysimproject_b200 import _lib as L
from nbodysimproject_b200 import ensemble as E

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 2 * 128 * 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
rng = np.random.RandomState(0)
m = rng.uniform(0.1, 10, (B, N)); q = rng.randn(B, N, 2) * 2.0; v = rng.randn(B, N, 2) * 0.3
bk = E.DeviceBucket(m, q, v, 0.05, 1.0, "yoshida4")
bk.n_sub[:] = 1
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bk.run(0.001, steps, 0, 0, flags=0, want_dyn=False)
    e1.record()
    torch.cuda.synchronize()
t = e0.elapsed_time(e1) * 1e-3
fl = B * steps * (3 * 14.0 * N * (N - 1) + 36.0 * N)
print(f"main kernel N={N} B={B} steps={steps}: {t*1e3:.2f} ms, {fl/t*1e-12:.2f} TFLOP/s algorithmic, "
      f"{B*steps/t:.3e} system-steps/s")
if len(sys.argv) > 4:
    from nbodysimproject_b200.largen import LargeNSimulation, make_disc
    n = int(sys.argv[4])
    mm, qq, vv = make_disc(n, 1)
    sim = LargeNSimulation(mm, qq, vv, softening=1e-3)
    for rep in range(3):
        e0.record(); sim.accelerations(); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    print(f"largeN n={n}: {t*1e3:.2f} ms, {float(n)*n/t:.3e} pairs/s, {14.0*n*n/t*1e-12:.2f} TFLOP/s(14/pair)")
