"""Short driver for ncu: the large-N force kernel.  python tools/profile_largen.py [n] [variant ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodysimproject_b200 import _lib as L
from nbodysimproject_b200.largen import LargeNSimulation, make_disc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
variants = [int(a) for a in sys.argv[2:]] or [-1]
mm, qq, vv = make_disc(n, 1)
sim = LargeNSimulation(mm, qq, vv, softening=1e-3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for var in variants:
    sim.variant = var
    for rep in range(3):
        e0.record(); sim.accelerations(); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    print(f"largeN n={n} variant={var}: {t*1e3:.2f} ms, {float(n)*n/t:.3e} pairs/s, {14.0*n*n/t*1e-12:.2f} TFLOP/s(14/pair)")
