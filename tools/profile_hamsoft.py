"""Short driver for ncu: hamsoft_run_kernel<3> on jittered README systems.  python tools/profile_hamsoft.py [B] [steps]"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from nbodysimproject_b200 import hamsoft as H

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 15
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
m, q, v = bench._c1_inputs(B, 42)
hs, s0 = H.default_params(object(), 1e-3, 1e-4, B)
ep = np.stack([np.maximum(s0, hs[:, H.P["eps_min"]]), np.zeros(B)], 1)
hb = H.HamSoftBucket(m, q, v, hs, ep, 1.0)
hb.setup(calibrate=True, freeze_dt=0.01)
print("n_sub", int(hb.n_sub.max()), int(hb.n_sub.min()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
work = torch.zeros((B, 2), dtype=torch.float64, device="cuda")
for rep in range(3):
    e0.record(); hb.run(0.01, steps, work=work); e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) * 1e-3
w = work.sum(0).cpu().numpy()
print(f"hamsoft N=3 B={B} steps={steps}: {t*1e3:.2f} ms, {B*steps/t:.3e} system-steps/s; "
      f"mean Jacobi sweeps per eps* evaluation {w[0] / (13 * w[1]):.2f}")

# generic N: random compact systems
for N in (3, 4, 6, 8):
    rng = np.random.RandomState(1)
    Bn = B // 4
    m = rng.uniform(0.5, 3.0, (Bn, N)); q = rng.randn(Bn, N, 2) * 0.8; v = rng.randn(Bn, N, 2) * 0.4
    v -= (m[:, :, None] * v).sum(1, keepdims=True) / m.sum(1)[:, None, None]
    hs, s0 = H.default_params(object(), 0.05, 0.005, Bn)
    hb = H.HamSoftBucket(m, q, v, hs, np.stack([s0, np.zeros(Bn)], 1), 1.0)
    hb.setup(calibrate=True, freeze_dt=0.01)
    hb.n_sub[:] = 1
    work = torch.zeros((Bn, 2), dtype=torch.float64, device="cuda")
    for rep in range(2):
        e0.record(); hb.run(0.01, steps, work=work); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    w = work.sum(0).cpu().numpy()
    print(f"hamsoft N={N} B={Bn} steps={steps} (n_sub forced to 1): {t*1e3:.2f} ms, {Bn*steps/t:.3e} system-substeps/s; "
          f"mean Jacobi sweeps per eps* evaluation {w[0] / ((4 * N + 1) * w[1]):.2f}")
