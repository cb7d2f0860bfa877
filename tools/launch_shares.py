"""Per-kernel-family shares of the kernel time in an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
  python tools/launch_shares.py gpurun_out/r2_launches_bench_step.csv > profiles/r2_launch_shares_bench_step.csv
The list covers the warm-up steps and the timed step of the same command (identical work per step), launches are
serialised and cold-cache under ncu: compare SHARES, not absolutes."""
import collections
import csv
import re
import sys

UNIT = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
idx = {h: i for i, h in enumerate(rows[start])}
fam = collections.OrderedDict()
for r in rows[start + 1:]:
    if r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void ", "", r[idx["Kernel Name"]])
    name = re.sub(r"[<(].*$", "", name)
    ms = float(r[idx["Metric Value"]].replace(",", "")) * UNIT.get(r[idx["Metric Unit"]], 1e-6)
    n, t = fam.get(name, (0, 0.0))
    fam[name] = (n + 1, t + ms)
tot = sum(t for _, t in fam.values())
print("kernel_family,launches,sum_ms,share_of_kernel_time")
for name, (n, t) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"{name},{n},{t:.3f},{t / tot:.4f}")
