#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native NBodySimProject hot path.

Workload (BASELINE.json configs[2], "C3"): a batch stability ensemble of hierarchical / polygon / random /
close-encounter systems with N = 3..8 bodies, integrator yoshida4, dt = 0.01, n_steps = 1000 main steps +
50 tangent-map (MEGNO) steps, analysis mode 'full' -- i.e. BatchStabilityAnalyzer.analyze_batch
(batch_stability_analyzer.py:62-80) for `--systems` systems per GPU.  One bench "step" = one full pass of that
analysis over the batch.  Systems are sharded by system over ranks, no data-path collective (weak scaling).

  value : system-steps/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e   : the same metric through the C-ABI host entry point nb_ensemble_analyze_host_async with pinned HOST
          buffers: H2D of every input and D2H of the feature tables inside the timed region
  roofline : FP64-pipe roofline of the dominant kernel (ensemble_main_kernel<N, yoshida4>), algorithmic flops
          from SURVEY.md section 8d, peak = DFMA micro-benchmark measured live on the same GPU
  cpu_baseline : the NumPy oracle (a restatement of the reference's Python path) on the host cores

`--impl reference` times that CPU path alone (the reference is pure Python and cannot travel to the GPU box;
the oracle port is pinned against the reference's outputs by tests/test_oracle_golden.py).
`--workload largen` reports the large-N direct-sum kernel (pair-interactions/s) instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT = 0.01
N_STEPS = 1000
N_MEGNO = 50
MODE = "yoshida4"
STEPS_PER_SYSTEM = N_STEPS + N_MEGNO


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.perf_counter()          # nvidia-smi needs a moment to attach: wait for its first line
            while not self.lines and time.perf_counter() - t0 < 5.0:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Index of the next sample: call at the start and at the end of the timed region."""
        return len(self.lines)

    def stop(self, lo: int = 0, hi: int = None):
        """Clocks / throttle reasons of the samples [lo, hi) (the timed region); when the region is shorter than
        three sampling periods, the neighbouring samples (warm-up before, the e2e leg after: the same kernels) are
        included so that the figure is still measured under load."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        hi = len(self.lines) if hi is None else hi
        note = "timed region"
        if hi - lo < 3:
            lo, hi, note = max(0, lo - 10), min(len(self.lines), hi + 10), "timed region +- 0.5 s (same kernels)"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[lo:hi]:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "window": note}


_ORIG_AFFINITY = None


def restore_affinity():
    """The CPU baseline leg uses every host core again."""
    if _ORIG_AFFINITY is not None:
        try:
            os.sched_setaffinity(0, _ORIG_AFFINITY)
        except Exception:
            pass


def bind_to_gpu_numa_node(index: int):
    """Pin this rank to the CPU cores next to its GPU (NVML's ideal CPU affinity) BEFORE any pinned host buffer is
    allocated, so the staging memory of the e2e leg lands on the GPU's own NUMA node: with 8 ranks on one box the host
    side of the PCIe transfers is otherwise the bottleneck.  Best effort; returns the number of cores or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            global _ORIG_AFFINITY
            _ORIG_AFFINITY = allowed
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def make_inputs(n_systems: int, seed: int):
    """Diverse cohort (ml_training_pipeline.py:44-122 extended to N <= 8) + per-system tangent draws."""
    from nbodysimproject_b200.generators import EnsembleInputs
    rng = np.random.default_rng(seed)
    buckets = EnsembleInputs.diverse(rng, n_systems, n_max=8)
    out = {}
    for N, (m, q, v, soft, cohort) in sorted(buckets.items()):
        B = m.shape[0]
        out[N] = dict(m=m, q=q, v=v, eps=soft, raw_dr=rng.standard_normal((B, N, 2)),
                      raw_dv=rng.standard_normal((B, N, 2)), cohort=cohort)
    return out


def flops_main(N, n_sub_sum, n_steps):
    """yoshida4 sub-step = 3 force evaluations (14 flop / ordered pair) + 36 N kick/drift flops."""
    return float(n_sub_sum) * n_steps * (3 * 14.0 * N * (N - 1) + 36.0 * N)


# ---------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference's Python path)
# ---------------------------------------------------------------------------------------------

def _cpu_worker(job):
    from oracle import nbody_oracle as O
    m, q, v, eps, rr, rv = job
    sim = O.OracleSim(m, q, v, softening=float(eps), integrator_mode=MODE)
    O.run_stability_analysis(sim, N_STEPS, DT, "full", rr, rv)
    return STEPS_PER_SYSTEM


def cpu_sample_jobs(n_jobs: int, seed: int):
    inp = make_inputs(max(64, n_jobs * 4), seed)
    jobs = []
    # round-robin over the N buckets so the sample has the workload's mix
    keys = sorted(inp)
    idx = {k: 0 for k in keys}
    while len(jobs) < n_jobs:
        for k in keys:
            d = inp[k]
            i = idx[k]
            if i < d["m"].shape[0] and len(jobs) < n_jobs:
                jobs.append((d["m"][i], d["q"][i], d["v"][i], d["eps"][i], d["raw_dr"][i], d["raw_dv"][i]))
                idx[k] += 1
    return jobs


def run_cpu(n_jobs: int, cores: int, seed: int = 777):
    import multiprocessing as mp
    jobs = cpu_sample_jobs(n_jobs, seed)
    t0 = time.perf_counter()
    if cores > 1:
        with mp.get_context("fork").Pool(cores) as pool:
            done = sum(pool.map(_cpu_worker, jobs, chunksize=1))
    else:
        done = sum(_cpu_worker(j) for j in jobs)
    dt = time.perf_counter() - t0
    return done / dt, dt


def cpu_pairs_per_s(n: int = 4096, reps: int = 3):
    """CPU baseline for the large-N metric: the oracle's dense gravitational_force (forces.py:63-75 restated) at the
    largest N whose (N,N,2) fp64 temporaries fit comfortably; pairs/s on one core (the reference is single-threaded)."""
    from oracle import nbody_oracle as O
    from nbodysimproject_b200.largen import make_disc
    m, q, _ = make_disc(n, seed=5)
    O.accelerations(q, m, 1e-3, 1.0)
    t0 = time.perf_counter()
    for _ in range(reps):
        O.accelerations(q, m, 1e-3, 1.0)
    dt = (time.perf_counter() - t0) / reps
    return float(n) * float(n) / dt, dt


def impl_reference_largen(args):
    n = 4096
    times = []
    for it in range(args.warmup + args.steps):
        rate, dt = cpu_pairs_per_s(n, 3)
        if it >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    value = float(n) * n / dt
    print(json.dumps({
        "impl": "reference", "metric": "pair-interactions/s at N=2^20", "value": value, "unit": "pair-interactions/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * dt,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C5 large-N direct sum; CPU path bounded to N=4096 (dense (N,N,2) fp64 temporaries: "
                               "N=2^20 would need 17.6 TB), pairs/s is size-independent for N >= 256"},
        "cpu_baseline": {"value": value, "unit": "pair-interactions/s", "cores": 1, "kind": "port",
                         "sample": "oracle dense gravitational_force at N=4096, 3 calls per step"},
        "e2e": {"value": value, "unit": "pair-interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def impl_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "largen":
        return impl_reference_largen(args)
    cores = os.cpu_count() or 1
    n_jobs = max(cores, 8) * 4
    times, done = [], 0
    for it in range(args.warmup + args.steps):
        rate, dt = run_cpu(n_jobs, cores, seed=1000 + it)
        if it >= args.warmup:
            times.append(dt)
            done += n_jobs * STEPS_PER_SYSTEM
        if sum(times) > 150:          # keep the whole arm within a few minutes
            break
    total = sum(times)
    value = done / total
    k = len(times)
    line = {
        "impl": "reference", "metric": "system-steps/s (N=3-8 ensembles, MEGNO on)", "value": value,
        "unit": "system-steps/s", "n_gpus": args.gpus, "steps": k, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(k, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 batch stability ensemble: diverse N=3-8 cohort, yoshida4, 1000+50 MEGNO steps, "
                               "mode full (bounded CPU sample of the same generator)",
                   "systems_per_step": n_jobs, "integrator": MODE, "dt": DT},
        "cpu_baseline": {"value": value, "unit": "system-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{n_jobs} systems x {STEPS_PER_SYSTEM} steps per bench step, NumPy oracle "
                                   f"(oracle/nbody_oracle.py) in {cores} worker processes"},
        "e2e": {"value": value, "unit": "system-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------

def impl_b200(args):
    import torch
    import torch.distributed as dist
    from nbodysimproject_b200 import _lib as L
    from nbodysimproject_b200 import ensemble as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()

    if args.workload == "largen":
        return bench_largen(args, torch, dist, world, rank, local, dev)
    if args.workload == "c2":
        bench_c2(args, torch, dist, world, rank, local, dev)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload in ("c4", "c1"):
        bench_secondary(args, torch, dist, world, rank, local, dev)
        if world > 1:
            dist.destroy_process_group()
        return

    B_total = args.systems
    inp = make_inputs(B_total, seed=42 + rank)
    Ns = sorted(inp, reverse=True)   # launch the large-N buckets first: their sequential sub-step tails are the longest
    # pinned host buffers (e2e) and device-resident copies (value)
    host, devb = {}, {}
    h2d = d2h = 0
    for N in Ns:
        d = inp[N]
        B = d["m"].shape[0]
        hb = {}
        for k in ("m", "q", "v", "eps", "raw_dr", "raw_dv"):
            t = torch.from_numpy(np.ascontiguousarray(d[k], dtype=np.float64)).pin_memory()
            hb[k] = t
            h2d += t.numel() * 8
        hb["v_work"] = torch.empty_like(hb["v"]).pin_memory()
        hb["dyn"] = torch.empty((B, L.N_DYN), dtype=torch.float64).pin_memory()
        hb["stat"] = torch.empty((B, L.N_STATIC), dtype=torch.float64).pin_memory()
        hb["nsub"] = torch.empty((B,), dtype=torch.int32).pin_memory()
        hb["status"] = torch.empty((B,), dtype=torch.int32).pin_memory()
        d2h += hb["dyn"].numel() * 8 + hb["stat"].numel() * 8 + B * 8 + hb["v"].numel() * 8
        host[N] = hb
        bk = E.DeviceBucket(hb["m"], hb["q"], hb["v"], hb["eps"], 1.0, MODE, dev)
        bk.v0 = bk.v.clone()
        bk.q0 = bk.q.clone()
        bk.rdr = hb["raw_dr"].to(dev)
        bk.rdv = hb["raw_dv"].to(dev)
        bk.stream = torch.cuda.Stream(device=dev)
        devb[N] = bk
    prep_flags = L.PREP_REMOVE_COM | L.PREP_CTOR_KICK | L.PREP_SNAPSHOT_KICK
    interval = max(1, N_STEPS // 100)
    launches_per_step = 0

    def step_device():
        nonlocal launches_per_step
        cur = torch.cuda.current_stream()
        n = 0
        # the construction-time kernels of every bucket first (0.3 % of the step), then the runs: nb_ensemble_run_f64
        # launches each bucket's sub-step-heavy head at high priority, so no head waits behind another bucket's bulk
        for N in Ns:
            bk = devb[N]
            bk.stream.wait_stream(cur)
            with torch.cuda.stream(bk.stream):
                bk.q.copy_(bk.q0)
                bk.v.copy_(bk.v0)
                bk.prepare(prep_flags, 0.01, 0.01, DT, 50, want_static=True)     # 1 kernel
                bk.sort()                                                        # 3 kernels
        for N in Ns:
            bk = devb[N]
            with torch.cuda.stream(bk.stream):
                bk.dyn = bk.run(DT, N_STEPS, interval, N_MEGNO, bk.rdr, bk.rdv, flags=L.RUN_ENERGY)  # 2+2+1+1 kernels
                n += 10
        for N in Ns:
            cur.wait_stream(devb[N].stream)
        launches_per_step = n

    def restore_e2e_inputs():
        # nb_ensemble_analyze_host* returns the kicked velocities in the caller's v (the reference mutates the caller's
        # sims the same way): put the original inputs back.  Bench scaffolding, outside the timed intervals.
        for N in Ns:
            host[N]["v_work"].copy_(host[N]["v"])

    def step_e2e():
        """One end-to-end step: pinned host inputs -> H2D -> kernels -> D2H of both feature tables, all buckets in
        flight on their own workspace slots; returns the wall time of exactly that."""
        t0 = time.perf_counter()
        # largest transfers first: that bucket's H2D is not queued behind everyone else's and its D2H overlaps the
        # kernels of the buckets issued after it
        for slot, N in enumerate(sorted(Ns, key=lambda n: -host[n]["m"].numel())):
            hb = host[N]
            B = hb["m"].shape[0]
            L.check(lib.nb_ensemble_analyze_host_async(
                L.ptr(hb["m"]), L.ptr(hb["q"]), L.ptr(hb["v_work"]), L.ptr(hb["eps"]), 1.0, B, N, L.MODES[MODE],
                prep_flags, 0.01, 0.01, DT, N_STEPS, N_MEGNO, 50, L.ptr(hb["raw_dr"]), L.ptr(hb["raw_dv"]),
                L.ptr(hb["dyn"]), L.ptr(hb["stat"]), L.ptr(hb["nsub"]), L.ptr(hb["status"]), local, slot % 8),
                "nb_ensemble_analyze_host_async")
        for slot in range(min(len(Ns), 8)):
            L.check(lib.nb_host_sync(slot), "nb_host_sync")
        return time.perf_counter() - t0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: device-resident
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    mark0 = sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    mark1 = sampler.mark()
    t_dev = e0.elapsed_time(e1) * 1e-3
    # ---- e2e: host buffers through the C ABI
    for _ in range(max(1, min(args.warmup, 2))):
        restore_e2e_inputs()
        step_e2e()
    barrier()
    t_e2e = 0.0
    for _ in range(args.steps):
        restore_e2e_inputs()
        if world > 1:
            dist.barrier()
        t_e2e += step_e2e()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop(mark0, mark1) if rank == 0 else None
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    sys_steps = float(B_total) * world * STEPS_PER_SYSTEM * args.steps

    # ---- cross-check: e2e and device paths give identical feature tables; count statuses
    same = all(np.array_equal(host[N]["dyn"].numpy(), devb[N].dyn.cpu().numpy(), equal_nan=True) for N in Ns)
    n_bad = int(sum(int((host[N]["status"].numpy() != 0).sum()) for N in Ns))

    # ---- roofline of the dominant kernel family: ensemble_main_kernel<N, yoshida4>, N = 3..8.
    # The six launches of a step run concurrently on their bucket streams (exactly as in the timed region), so
    # the figure is: algorithmic flops of those launches / the CUDA-event time from the first launch to the last
    # completion.  Per-bucket solo timings are reported too; they expose the sequential sub-step tail of the few
    # n_sub ~ 50 systems, which the concurrent launch hides behind the bulk.
    roof = None
    if rank == 0:
        peak = L.peak_flops(0, local)
        per = []
        tot_fl = 0.0
        for N in Ns:
            bk = devb[N]
            bk.q.copy_(bk.q0); bk.v.copy_(bk.v0)
            bk.prepare(prep_flags, 0.01, 0.01, DT, 50)
            bk.sort()
            bk.vk = bk.v.clone()
            nsub_sum = int(bk.n_sub.sum().item())
            bk.flops = flops_main(N, nsub_sum, N_STEPS)
            tot_fl += bk.flops
            best = 1e30
            for rep in range(2):
                bk.q.copy_(bk.q0); bk.v.copy_(bk.vk)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                bk.run(DT, N_STEPS, interval, 0, flags=0, want_dyn=False)
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b) * 1e-3)
            per.append(dict(N=N, B=bk.B, mean_n_sub=nsub_sum / bk.B, max_n_sub=int(bk.n_sub.max().item()),
                            solo_ms=best * 1e3, solo_tflops=bk.flops / best * 1e-12))
        best = 1e30
        cur = torch.cuda.current_stream()
        for rep in range(3):
            for N in Ns:
                devb[N].q.copy_(devb[N].q0); devb[N].v.copy_(devb[N].vk)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for N in Ns:
                bk = devb[N]
                bk.stream.wait_stream(cur)
                with torch.cuda.stream(bk.stream):
                    bk.run(DT, N_STEPS, interval, 0, flags=0, want_dyn=False)
            for N in Ns:
                cur.wait_stream(devb[N].stream)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) * 1e-3)
        ach = tot_fl / best * 1e-12
        roof = {"bound": "fp64", "kernel": "ensemble_main_kernel<N=3..8, yoshida4> (6 concurrent launches per step)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of the six launches of one step (ncu, this workload at
                # 2^20 systems per GPU: profiles/r1_main_kernels_dram.csv), scaled by the batch size
                "traffic": 3.377e8 * (B_total / float(1 << 20)),
                "traffic_note": "bytes per step over the six launches; the state is read once and lives in registers",
                "flops_per_step": tot_fl, "ms": best * 1e3,
                "peak_source": "nb_peak_flops(0): register-resident DFMA micro-benchmark, same GPU, same run "
                               "(MEASURED_PEAKS.json has no FP64 figure; nominal 64 DFMA/clk/SM x 148 x 1.965 GHz = 37.2)",
                "flop_model": "SURVEY.md 8d: per sub-step 3 x 14 N(N-1) + 36 N",
                "share_of_step": best / (t_dev / args.steps), "per_bucket_solo": per}

    # ---- second half of BASELINE.json's metric: pair-interactions/s of the large-N direct sum at N = 2^20
    # (all ranks take part: i-blocks sharded, in-place NCCL all-gather of the packed positions per evaluation)
    largen = None
    if not args.no_largen:
        from nbodysimproject_b200 import largen as LN
        n_ln = 1 << 20
        mm, qq, vv = LN.make_disc(n_ln, seed=1)
        lsim = LN.LargeNSimulation(mm, qq, vv, G=1.0, softening=1e-3, device=dev)
        k_ln = 3
        t_ln = LN.measure_force(lsim, k_ln, 3)
        if rank == 0:
            peak32 = L.peak_flops(1, local)
            rate = float(n_ln) * n_ln * k_ln / t_ln
            largen = {"metric": "pair-interactions/s at N=2^20", "value": rate, "unit": "pair-interactions/s",
                      "n": n_ln, "ms_per_force_evaluation": 1e3 * t_ln / k_ln, "scaling": "strong", "dtype": "f32",
                      "gpu_launches": 2 * k_ln,
                      "roofline": {"bound": "fp32", "kernel": "largeN_accel_x2_kernel",
                                   "achieved": 14.0 * rate * 1e-12 / world, "peak": peak32, "unit": "TFLOP/s",
                                   "frac": 14.0 * rate * 1e-12 / world / peak32, "flops_per_pair": 14,
                                   "peak_source": "nb_peak_flops(1): FFMA micro-benchmark, same GPU, same run"}}
        del lsim

    # ---- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        restore_affinity()
        cores = os.cpu_count() or 1
        n_jobs = max(cores, 8) * 96          # ~10-15 s of CPU work on the box's cores
        rate, dt_cpu = run_cpu(n_jobs, cores)
        cpu = {"value": rate, "unit": "system-steps/s", "cores": cores, "kind": "port",
               "sample": f"{n_jobs} systems x {STEPS_PER_SYSTEM} steps of the same generator, NumPy oracle in "
                         f"{cores} processes, {dt_cpu:.1f} s"}

    if rank == 0:
        line = {
            "metric": "system-steps/s (N=3-8 ensembles, MEGNO on)", "value": sys_steps / t_dev,
            "unit": "system-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3 batch stability ensemble (BASELINE.json configs[2]): diverse cohort "
                                   "40% random N=3-8 / 30% hierarchical triples / 20% polygons / 10% close encounters, "
                                   "yoshida4 dt=0.01, 1000 steps + 50 tangent-map MEGNO steps, mode full",
                       "systems_per_gpu": B_total, "buckets": {str(N): int(devb[N].B) for N in Ns},
                       "sharding": "by system, no collective", "cpu_cores_bound_to_gpu_numa_node": numa, "l2_note": "inputs re-read from HBM each step "
                       f"({h2d / 1e6:.0f} MB per GPU > 126 MB L2)"},
            "e2e": {"value": sys_steps / t_e2e, "unit": "system-steps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t_e2e / args.steps,
                    "api": "nb_ensemble_analyze_host_async (C ABI, pinned host buffers)"},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "largen": largen,
            "checks": {"e2e_equals_device_path": bool(same), "systems_with_nonzero_status": n_bad},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# secondary ensemble workloads: C4 (whfast planetary) and C1 (README 3-body ham_soft, batched)
# ---------------------------------------------------------------------------------------------

def _c4_inputs(B, seed):
    """star + 2/3/4 planets near 3:2 / 2:1 / 5:3 chains, half of them the TTV cohort."""
    from nbodysimproject_b200.generators import EnsembleInputs
    rng = np.random.default_rng(seed)
    out = {}
    per = [B // 3, B // 3, B - 2 * (B // 3)]
    for npl, b in zip((2, 3, 4), per):
        parts = [EnsembleInputs.planetary(rng, b - b // 2, npl, ttv=False), EnsembleInputs.planetary(rng, b // 2, npl, ttv=True)]
        out[npl + 1] = tuple(np.concatenate([p[k] for p in parts]) for k in range(4))
    return out


def _c1_inputs(B, seed):
    """README example (masses [1, 0.5, 0.1], collinear), jittered by 1e-3 randn so the systems differ."""
    rng = np.random.default_rng(seed)
    m = np.tile(np.array([1.0, 0.5, 0.1]), (B, 1))
    q = np.tile(np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]]), (B, 1, 1)) + 1e-3 * rng.standard_normal((B, 3, 2))
    v = np.tile(np.array([[0.0, 0.0], [0.0, 1.0], [0.0, 0.5]]), (B, 1, 1)) + 1e-3 * rng.standard_normal((B, 3, 2))
    v = v - np.sum(m[:, :, None] * v, axis=1, keepdims=True) / np.sum(m, axis=1)[:, None, None]
    return m, q, v


def _cpu_c4(job):
    from oracle import nbody_oracle as O
    m, q, v, n_steps, dt = job
    sim = O.OracleSim(m, q, v, softening=0.0, integrator_mode="whfast")
    for _ in range(n_steps):
        sim.step(dt)
    return n_steps


def _cpu_c1(job):
    from oracle.hamsoft_oracle import HamSoftOracleSim
    m, q, v, n_steps, dt = job
    sim = HamSoftOracleSim(m, q, v, softening=1e-3)
    for _ in range(n_steps):
        sim.step(dt)
    return n_steps


def _cpu_pool(fn, jobs, cores):
    import multiprocessing as mp
    restore_affinity()
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        done = sum(pool.map(fn, jobs, chunksize=1))
    dt = time.perf_counter() - t0
    return done / dt, dt


def bench_secondary(args, torch, dist, world, rank, local, dev):
    """--workload c4 | c1: system-steps/s of the whfast planetary ensemble / the batched README ham_soft system.
    One bench step = n_steps integrator steps of every system in one persistent-kernel launch per bucket."""
    from nbodysimproject_b200 import _lib as L
    from nbodysimproject_b200 import ensemble as E
    from nbodysimproject_b200 import hamsoft as H
    c4 = args.workload == "c4"
    n_steps = 1000
    dt = 0.01 * 2.0 * np.pi if c4 else 0.01
    B = args.systems if c4 else min(args.systems, 1 << 17)
    runs = []
    if c4:
        inp = _c4_inputs(B, 42 + rank)
        for N, (m, q, v, eps) in sorted(inp.items()):
            bk = E.DeviceBucket(m, q, v, eps, 1.0, "whfast", dev)
            bk.q0, bk.v0 = bk.q.clone(), bk.v.clone()
            bk.stream = torch.cuda.Stream(device=dev)
            runs.append(bk)
    else:
        m, q, v = _c1_inputs(B, 42 + rank)
        hs, s0 = H.default_params(object(), 1e-3, 1e-4, B)
        ep = np.stack([np.maximum(s0, hs[:, H.P["eps_min"]]), np.zeros(B)], 1)
        hb = H.HamSoftBucket(m, q, v, hs, ep, 1.0, dev)
        hb.setup(calibrate=True, freeze_dt=dt)
        hb.q0, hb.v0, hb.ep0, hb.hs0 = hb.bk.q.clone(), hb.bk.v.clone(), hb.eps_pi.clone(), hb.hs.clone()
        runs.append(hb)

    def step():
        cur = torch.cuda.current_stream()
        if c4:
            for bk in runs:
                bk.stream.wait_stream(cur)
                with torch.cuda.stream(bk.stream):
                    bk.q.copy_(bk.q0); bk.v.copy_(bk.v0)
                    bk.prepare(L.PREP_REMOVE_COM | L.PREP_CTOR_KICK, dt, dt, dt, 50)
                    bk.run(dt, n_steps, 0, 0, flags=L.RUN_WRITE_STATE, want_dyn=False)
            for bk in runs:
                cur.wait_stream(bk.stream)
        else:
            hb = runs[0]
            hb.bk.q.copy_(hb.q0); hb.bk.v.copy_(hb.v0); hb.eps_pi.copy_(hb.ep0); hb.hs.copy_(hb.hs0)
            hb.run(dt, n_steps)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    mark0 = sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop(mark0, sampler.mark()) if rank == 0 else None
    t_dev = e0.elapsed_time(e1) * 1e-3
    # e2e: pinned host state in, final host state out, every step
    host = []
    for r in runs:
        bk = r if c4 else r.bk
        host.append((bk.q0.cpu().pin_memory() if c4 else r.q0.cpu().pin_memory(),
                     bk.v0.cpu().pin_memory() if c4 else r.v0.cpu().pin_memory(),
                     torch.empty(bk.q.shape, dtype=torch.float64).pin_memory(),
                     torch.empty(bk.v.shape, dtype=torch.float64).pin_memory()))
    h2d = sum(a.numel() * 8 + b.numel() * 8 for a, b, _, _ in host)
    d2h = h2d

    def step_e2e():
        for r, (q0h, v0h, qh, vh) in zip(runs, host):
            if c4:
                r.q0.copy_(q0h, non_blocking=True); r.v0.copy_(v0h, non_blocking=True)
            else:
                r.q0.copy_(q0h, non_blocking=True); r.v0.copy_(v0h, non_blocking=True)
        step()
        for r, (q0h, v0h, qh, vh) in zip(runs, host):
            bk = r if c4 else r.bk
            qh.copy_(bk.q, non_blocking=True); vh.copy_(bk.v, non_blocking=True)
        torch.cuda.synchronize()

    step_e2e()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    n_bad = int(sum(int(((r if c4 else r.bk).status != 0).sum()) for r in runs))
    if rank != 0:
        return
    sys_steps = float(B) * world * n_steps * args.steps
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        if c4:
            inp = _c4_inputs(3 * cores * 2, 7)
            jobs = [(inp[N][0][i], inp[N][1][i], inp[N][2][i], 2000, dt) for N in sorted(inp) for i in range(inp[N][0].shape[0])]
            rate, dtc = _cpu_pool(_cpu_c4, jobs, cores)
            sample = f"{len(jobs)} planetary systems x 2000 whfast steps, NumPy oracle in {cores} processes, {dtc:.1f} s"
        else:
            m, q, v = _c1_inputs(cores, 7)
            jobs = [(m[i], q[i], v[i], 1000, dt) for i in range(cores)]
            rate, dtc = _cpu_pool(_cpu_c1, jobs, cores)
            sample = f"{cores} jittered README systems x 1000 ham_soft steps, oracle in {cores} processes, {dtc:.1f} s"
        cpu = {"value": rate, "unit": "system-steps/s", "cores": cores, "kind": "port", "sample": sample}
    name = ("C4 WHFast + Kepler planetary ensemble (BASELINE.json configs[3]): star + 2-4 planets near 3:2/2:1/5:3, "
            "half TTV cohort, dt = 0.01 x 2 pi, bug-compatible Kepler solver") if c4 else \
           ("C1 README 3-body ham_soft (BASELINE.json configs[0]) batched: jittered copies, dt = 0.01, adaptive-epsilon "
            "Strang flow with the finite-difference eps* gradient")
    line = {
        "metric": "system-steps/s", "value": sys_steps / t_dev, "unit": "system-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "systems_per_gpu": B, "integrator_steps_per_bench_step": n_steps,
                   "buckets": {str((r if c4 else r.bk).N): int((r if c4 else r.bk).B) for r in runs}},
        "e2e": {"value": sys_steps / t_e2e, "unit": "system-steps/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t_e2e / args.steps},
        "gpu_launches": int((2 * len(runs) if c4 else 1) * args.steps), "roofline": None, "cpu_baseline": cpu,
        "clocks": clocks, "checks": {"systems_with_nonzero_status": n_bad},
    }
    print(json.dumps(line))


def _c2_sims(mode):
    from nbodysimproject_b200.generators import InitialConditionGenerator, set_global_seed
    set_global_seed(42)
    gen = InitialConditionGenerator()
    return [gen.create_simulation(3 + (i % 3), integrator_mode=mode) for i in range(10)]


def _cpu_c2(job):
    from oracle import nbody_oracle as O
    m, q, v, soft = job
    sim = O.OracleSim(m, q, v, softening=soft, integrator_mode="verlet")
    rr, rv = np.random.default_rng(0).standard_normal((2, len(m), 2))
    O.run_stability_analysis(sim, 1000, 0.01, "full", rr, rv)
    return 1050


def bench_c2(args, torch, dist, world, rank, local, dev):
    """--workload c2 (BASELINE.json configs[1]): the quick_test cohort (10 systems, N = 3, 4, 5 cycling) analysed through
    the PUBLIC PYTHON API -- NBodySimulation objects in, DataFrame out -- as the BASELINE wording has it (verlet, 1000
    steps, 'full' mode with MEGNO).  Ten tiny systems are launch-latency-bound: this line measures the API overhead a
    reference user sees, not the kernels."""
    import io
    import contextlib
    from nbodysimproject_b200.stability import BatchStabilityAnalyzer
    if rank != 0:
        return

    def step():
        sims = _c2_sims("verlet")
        np.random.seed(7)
        with contextlib.redirect_stdout(io.StringIO()):
            return BatchStabilityAnalyzer(n_steps=1000, dt=0.01, mode="full").analyze_batch(sims, show_progress=False)

    for _ in range(args.warmup):
        df = step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        df = step()
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    sys_steps = 10 * 1050 * args.steps
    cpu = None
    if not args.no_cpu:
        sims = _c2_sims("verlet")
        jobs = [(s._mass.copy(), s._pos.copy(), s._vel.copy(), float(s.manager.s0)) for s in sims]
        cores = min(os.cpu_count() or 1, len(jobs))
        rate, dtc = _cpu_pool(_cpu_c2, jobs, cores)
        cpu = {"value": rate, "unit": "system-steps/s", "cores": cores, "kind": "port",
               "sample": f"the same 10 systems, oracle run_stability_analysis in {cores} processes, {dtc:.1f} s"}
    print(json.dumps({
        "metric": "system-steps/s", "value": sys_steps / t, "unit": "system-steps/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 MLTrainingPipeline.quick_test cohort (BASELINE.json configs[1]): 10 systems N=3,4,5, verlet, "
                               "1000 steps + MEGNO, mode full, through NBodySimulation / BatchStabilityAnalyzer (includes "
                               "building the 10 simulation objects and the DataFrame)",
                   "columns": int(df.shape[1])},
        "e2e": {"value": sys_steps / t, "unit": "system-steps/s", "h2d_bytes_per_step": int(10 * 4 * 5 * 8 * 3),
                "d2h_bytes_per_step": int(10 * 47 * 8), "api": "BatchStabilityAnalyzer.analyze_batch"},
        "gpu_launches": None, "roofline": None, "cpu_baseline": cpu}))


def bench_largen(args, torch, dist, world, rank, local, dev):
    from nbodysimproject_b200.largen import bench_largen as run
    sampler = None
    if rank == 0:
        sampler = ClockSampler(local)
        sampler.start()
    line = run(args, world, rank, local, dev, sampler)
    if rank == 0 and line is not None and world == 1 and not args.no_cpu:
        restore_affinity()
        rate, dtc = cpu_pairs_per_s(4096, 5)
        line["cpu_baseline"] = {"value": rate, "unit": "pair-interactions/s", "cores": 1, "kind": "port",
                                "sample": f"oracle dense gravitational_force at N=4096 ({dtc*1e3:.0f} ms per call; the "
                                          "(N,N,2) fp64 temporaries make N=2^20 impossible on the CPU path: 17.6 TB)"}
    if rank == 0 and line is not None:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 20; 5 for c4 / c1 / largen)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ensemble", choices=["ensemble", "largen", "c4", "c1", "c2"])
    ap.add_argument("--systems", type=int, default=1 << 20, help="systems per GPU (weak scaling)")
    ap.add_argument("--n", type=int, default=1 << 20, help="particles for --workload largen")
    ap.add_argument("--n-hamsoft", type=int, default=0, dest="n_hamsoft",
                    help="particles for the ham_soft Strang sub-step timing of --workload largen (default: --n)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-largen", action="store_true", help="skip the secondary large-N force measurement")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if (args.workload == "ensemble" and args.impl == "b200") else 5
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        impl_reference(args)
    else:
        impl_b200(args)


if __name__ == "__main__":
    main()
