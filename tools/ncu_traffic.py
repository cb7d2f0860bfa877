"""profiles/r2_traffic.json: DRAM bytes per launch of the dominant kernels, from ncu captures of the bench commands
(dram__bytes_read.sum + dram__bytes_write.sum).  bench.py reads it for `roofline.traffic`.
  python tools/ncu_traffic.py gpurun_out/r2_main_dram.csv [gpurun_out/r2_c1_dram.csv gpurun_out/r2_c4_dram.csv gpurun_out/r2_largen_dram.csv]
Each CSV is the --csv log of `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
--kernel-name regex:<kernel> ... python bench.py --steps 1 --warmup 3 ...` (the LAST step's launches are the timed ones)."""
import collections
import csv
import json
import os
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    idx = {h: i for i, h in enumerate(rows[start])}
    data = collections.OrderedDict()
    for r in rows[start + 1:]:
        key = (int(r[idx["ID"]]), r[idx["Kernel Name"]])
        val = float(r[idx["Metric Value"]].replace(",", "")) * UNIT.get(r[idx["Metric Unit"]], 1.0)
        data.setdefault(key, {})[r[idx["Metric Name"]]] = val
    return data


def total(d, keys):
    return float(sum(d[k].get("dram__bytes_read.sum", 0.0) + d[k].get("dram__bytes_write.sum", 0.0) for k in keys))


def main():
    out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_traffic.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for path in sys.argv[1:]:
        d = launches(path)
        keys = list(d.keys())
        name = os.path.basename(path)
        if "main" in name:
            # one step = 6 buckets x (head launch + rest launch): the last 12 captured launches are the timed step
            last = keys[-12:]
            out["ensemble_main_six_launches_bytes_per_2p20_systems"] = total(d, last)
            out["ensemble_main_detail"] = [{"kernel": k[1], "bytes": total(d, [k]), "ns": d[k].get("gpu__time_duration.sum")} for k in last]
        elif "c1" in name:
            out["c1_hamsoft_bytes_per_launch"] = total(d, keys[-1:])
        elif "c4" in name:
            out["c4_main_bytes_per_step"] = total(d, keys[-3:])          # three buckets (N = 3, 4, 5) per step
        elif "largen" in name:
            out["largeN_accel_bytes_per_launch_2p20"] = total(d, keys[-1:])
        out.setdefault("sources", {})[name] = "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (cold-cache, serialised launches)"
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if not isinstance(v, (list, dict))}, indent=1))


if __name__ == "__main__":
    main()
