"""On-GPU micro-benchmarks used while tuning (not part of the product path).
python tools/microbench.py [peaks] [ensemble] [largen]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodysimproject_b200 import _lib as L
from nbodysimproject_b200 import ensemble as E


def ev_time(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def peaks():
    for w, name in enumerate(["fp64 DFMA TFLOP/s", "fp32 FFMA TFLOP/s", "fp32x2 FFMA2 TFLOP/s", "MUFU.RSQ f32 Tops/s",
                              "MUFU.RSQ64H Tops/s"]):
        print(name, round(L.peak_flops(w), 3), flush=True)


def ensemble(B=1 << 17, steps=200):
    rng = np.random.RandomState(0)
    for mode in ("verlet", "yoshida4"):
        for N in range(3, 9):
            m = rng.uniform(0.1, 10, (B, N))
            q = rng.randn(B, N, 2) * 2.0
            v = rng.randn(B, N, 2) * 0.3
            bk = E.DeviceBucket(m, q, v, 0.05, 1.0, mode)
            bk.n_sub[:] = 1
            t = ev_time(lambda: bk.run(0.001, steps, 0, 0, flags=0, want_dyn=False))
            evals = (1 if mode == "verlet" else 3) * steps * B
            flops = evals * 14.0 * N * (N - 1)
            print(f"{mode} N={N} B={B}: {B*steps/t:.3e} system-steps/s  {flops/t*1e-12:.2f} TFLOP/s (14 flop/pair) "
                  f"{evals*N*(N-1)/t:.3e} pairs/s", flush=True)


def largen(n=1 << 18):
    """All kernel variants of the large-N force (the variant is an argument of nb_largeN_accel_f32)."""
    import ctypes
    rng = np.random.RandomState(0)
    xym = np.zeros((n, 4), dtype=np.float32)
    xym[:, :2] = rng.randn(n, 2)
    xym[:, 2] = rng.uniform(0.5, 1.5, n) / n
    d = torch.as_tensor(xym).cuda()
    acc = torch.empty((n, 2), dtype=torch.float32, device="cuda")
    sums = torch.zeros(2, dtype=torch.float64, device="cuda")
    ws = torch.empty((n, 2), dtype=torch.float64, device="cuda")
    lib = L.load()
    for variant in (1, 8, 9, 10):
      for with_sums in (False, True):
          def run():
              L.check(lib.nb_largeN_accel_f32(L.ptr(d), n, 0, n, 1e-3, 1.0, L.ptr(acc),
                                              L.ptr(sums) if with_sums else None, L.ptr(ws), variant, L.stream_ptr()))
          run()
          t = ev_time(run)
          pairs = float(n) * n
          print(f"largeN n={n} variant={variant} sums={with_sums}: {pairs/t:.3e} pairs/s "
                f"{pairs*14/t*1e-12:.2f} TFLOP/s(14/pair) t={t*1e3:.2f} ms", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["peaks", "ensemble", "largen"]
    if "peaks" in which:
        peaks()
    if "ensemble" in which:
        ensemble()
    if "largen" in which:
        largen()
