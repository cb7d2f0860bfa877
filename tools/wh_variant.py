"""Tuning helper: time the whfast main kernel of one library build on the C4 cohort and print a hash of the final state
(bit-identity across builds).  usage: python tools/wh_variant.py path/to/lib.so [B]"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodysimproject_b200 import _lib as L  # noqa: E402

L.LIB_PATH = os.path.abspath(sys.argv[1])
import torch  # noqa: E402
import bench  # noqa: E402
from nbodysimproject_b200 import ensemble as E  # noqa: E402

B = int(sys.argv[2]) if len(sys.argv) > 2 else 3 * 131072
DT = 0.01 * 2.0 * np.pi
inp = bench._c4_inputs(B, 5)
tot_t, tot_steps = 0.0, 0
for N, (m, q, v, eps) in sorted(inp.items()):
    bk = E.DeviceBucket(m, q, v, eps, 1.0, "whfast")
    bk.prepare(L.PREP_REMOVE_COM | L.PREP_CTOR_KICK, DT, DT, DT, 50)
    q0, v0 = bk.q.clone(), bk.v.clone()
    best = 1e9
    for rep in range(3):
        bk.q.copy_(q0); bk.v.copy_(v0)
        work = torch.zeros((bk.B, 2), dtype=torch.float64, device=bk.device)
        torch.cuda.synchronize()
        t = time.perf_counter()
        bk.run(DT, 1000, 0, 0, flags=L.RUN_WRITE_STATE, want_dyn=False, work=work)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    h = hashlib.sha256(bk.q.cpu().numpy().tobytes() + bk.v.cpu().numpy().tobytes()).hexdigest()[:16]
    w = work.cpu().numpy()
    print(f"N={N} B={bk.B} {best * 1e3:8.2f} ms  {bk.B * 1000 / best:.3e} system-steps/s  iters/solve {w[:, 0].sum() / w[:, 1].sum():.3f}  "
          f"status!=0 {int((bk.status != 0).sum())}  state {h}")
    tot_t += best; tot_steps += bk.B * 1000
print(f"all: {tot_steps / tot_t:.3e} system-steps/s (buckets run one after the other)")
