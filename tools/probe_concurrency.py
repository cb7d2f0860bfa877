"""When does each bucket's main kernel finish inside the concurrent step?  python tools/probe_concurrency.py [order]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from nbodysimproject_b200 import _lib as L, ensemble as E

order = sys.argv[1] if len(sys.argv) > 1 else "desc"
inp = bench.make_inputs(1 << 20, 42)
Ns = sorted(inp, reverse=(order == "desc"))
dev = torch.device("cuda", 0)
bks = {}
flags = L.PREP_REMOVE_COM | L.PREP_CTOR_KICK | L.PREP_SNAPSHOT_KICK
for N in Ns:
    d = inp[N]
    bk = E.DeviceBucket(d["m"], d["q"], d["v"], d["eps"], 1.0, "yoshida4", dev)
    bk.q0, bk.v0 = bk.q.clone(), bk.v.clone()
    bk.prepare(flags, 0.01, 0.01, 0.01, 50); bk.sort(); bk.vk = bk.v.clone()
    bk.stream = torch.cuda.Stream(device=dev)
    bks[N] = bk
cur = torch.cuda.current_stream()
for rep in range(3):
    for N in Ns:
        bks[N].q.copy_(bks[N].q0); bks[N].v.copy_(bks[N].vk)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); a.record()
    ends = {}
    for N in Ns:
        bk = bks[N]
        bk.stream.wait_stream(cur)
        with torch.cuda.stream(bk.stream):
            bk.run(0.01, 1000, 10, 0, flags=0, want_dyn=False)
            e = torch.cuda.Event(enable_timing=True); e.record(); ends[N] = e
    torch.cuda.synchronize()
    print(order, "rep", rep, {N: round(a.elapsed_time(ends[N]), 1) for N in Ns},
          "n_heavy/thr", {N: (int(bks[N]._bins[64]), int(bks[N]._bins[65])) for N in Ns})
