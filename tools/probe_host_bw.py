"""Pinned host <-> device copy bandwidth with all ranks copying at once (the floor under the e2e leg of bench.py).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/probe_host_bw.py
Each rank: H2D alone, D2H alone, both directions at once -- first with every rank active (what the e2e leg sees), then
rank 0 alone (what one GPU sees).  Prints one JSON line per configuration on rank 0 (min / mean over ranks, GB/s)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (NUMA binding helper)


def measure(dev, h_in, d_in, d_out, h_out, mode, reps=5):
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    nbytes = h_in.numel() * h_in.element_size() * (2 if mode == "both" else 1)
    return nbytes / best * 1e-9


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bench.bind_to_gpu_numa_node(local) if "--no-bind" not in sys.argv else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 512 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    for who in ("all_ranks", "rank0_alone"):
        for mode in ("h2d", "d2h", "both"):
            if world > 1:
                dist.barrier()
            bw = measure(dev, h_in, d_in, d_out, h_out, mode) if (who == "all_ranks" or rank == 0) else 0.0
            if world > 1:
                t = torch.tensor([bw], dtype=torch.float64, device=dev)
                lst = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(lst, t)
                vals = [float(x) for x in lst]
            else:
                vals = [bw]
            if rank == 0:
                act = vals if who == "all_ranks" else vals[:1]
                print(json.dumps({"probe": "pinned host<->device copy", "ranks_active": len(act), "direction": mode,
                                  "GBps_min": min(act), "GBps_mean": sum(act) / len(act), "GBps_sum": sum(act),
                                  "bytes_per_direction": n, "numa_bound_cores": numa}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
