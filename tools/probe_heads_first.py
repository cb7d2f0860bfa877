"""Experiment: does launching every bucket's sub-step-heavy head first (high-priority streams) shorten the concurrent
main-kernel phase?  python tools/probe_heads_first.py [head_fraction]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from nbodysimproject_b200 import _lib as L, ensemble as E

frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
inp = bench.make_inputs(1 << 20, 42)
Ns = sorted(inp, reverse=True)
dev = torch.device("cuda", 0)
flags = L.PREP_REMOVE_COM | L.PREP_CTOR_KICK | L.PREP_SNAPSHOT_KICK
whole, heads, rests = {}, {}, {}
lo_pri, hi_pri = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
for N in Ns:
    d = inp[N]
    bk = E.DeviceBucket(d["m"], d["q"], d["v"], d["eps"], 1.0, "yoshida4", dev)
    bk.prepare(flags, 0.01, 0.01, 0.01, 50); bk.sort()
    ns = bk.n_sub.cpu().numpy()
    order = np.argsort(-ns, kind="stable")
    nh = max(128, int(frac * len(ns)))
    vk = bk.v.cpu().numpy()
    for name, idx, store, pri in (("head", order[:nh], heads, -1), ("rest", order[nh:], rests, 0)):
        b = E.DeviceBucket(d["m"][idx], d["q"][idx], vk[idx], d["eps"][idx], 1.0, "yoshida4", dev)
        b.n_sub = bk.n_sub[torch.as_tensor(idx).to(dev)].contiguous()
        b.sort()
        b.q0, b.v0 = b.q.clone(), b.v.clone()
        b.stream = torch.cuda.Stream(device=dev, priority=pri)
        store[N] = b
    bk.q0, bk.vk = bk.q.clone(), bk.v.clone()
    bk.stream = torch.cuda.Stream(device=dev)
    whole[N] = bk
cur = torch.cuda.current_stream()

def timed(groups):
    best = 1e9
    for rep in range(3):
        for g in groups:
            for b in g.values():
                b.q.copy_(b.q0); b.v.copy_(b.vk if hasattr(b, "vk") else b.v0)
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for g in groups:
            for N in Ns:
                b = g[N]
                b.stream.wait_stream(cur)
                with torch.cuda.stream(b.stream):
                    b.run(0.01, 1000, 10, 0, flags=0, want_dyn=False)
        for g in groups:
            for N in Ns:
                cur.wait_stream(g[N].stream)
        e.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(e))
    return best

print("whole buckets            :", round(timed([whole]), 1), "ms")
print(f"heads ({frac:.0%}) first, hi-pri:", round(timed([heads, rests]), 1), "ms")
