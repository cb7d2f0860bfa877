"""Short driver for ncu: ensemble_main_kernel<4, whfast> on the C4 planetary cohort.  python tools/profile_whfast.py [B] [steps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from nbodysimproject_b200 import _lib as L, ensemble as E

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
inp = bench._c4_inputs(3 * B, 42)
dt = 0.01 * 2 * np.pi
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for N in (4,):
    m, q, v, eps = inp[N]
    bk = E.DeviceBucket(m, q, v, eps, 1.0, "whfast")
    bk.prepare(L.PREP_REMOVE_COM | L.PREP_CTOR_KICK, dt, dt, dt, 50)
    for rep in range(3):
        e0.record(); bk.run(dt, steps, 0, 0, flags=L.RUN_WRITE_STATE, want_dyn=False); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    print(f"whfast N={N} B={bk.B} steps={steps}: {t*1e3:.2f} ms, {bk.B*steps/t:.3e} system-steps/s")
