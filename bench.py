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
import math
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


def flops_megno(N, B, n_sub_sum, n_megno):
    """SURVEY.md section 8d, "system-step, yoshida4 + MEGNO" = 3 x 14 N(N-1) + 19 N(N-1) + 44 N: a MEGNO step is a
    macro step (n_sub sub-steps) plus one tangent-map evaluation (19 flop / ordered pair + 8 N for the tangent update)."""
    return flops_main(N, n_sub_sum, n_megno) + float(B) * n_megno * (19.0 * N * (N - 1) + 8.0 * N)


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own Python path when a checkout is importable (baseline/_ref/minbody or $NBODY_REFERENCE;
# the reference is pure Python, is not pip-installable and does not travel to the GPU box, so normally it is not),
# else the NumPy oracle port (pinned against the reference's outputs by tests/test_oracle_golden.py).
# ---------------------------------------------------------------------------------------------

_REAL_REF = None


def _real_reference():
    """Import the unmodified reference package if one is present; returns the module or None."""
    global _REAL_REF
    if _REAL_REF is not None:
        return _REAL_REF or None
    _REAL_REF = False
    cands = [os.path.join(ROOT, "baseline", "_ref")]
    if os.environ.get("NBODY_REFERENCE"):
        cands.insert(0, os.environ["NBODY_REFERENCE"])
    for c in cands:
        if os.path.isdir(os.path.join(c, "minbody")):
            import types
            sys.dont_write_bytecode = True
            sys.modules.setdefault("lightgbm", types.ModuleType("lightgbm"))   # minbody/__init__.py imports its trainer
            sys.path.insert(0, c)
            try:
                import minbody  # noqa: F401
                _REAL_REF = minbody
            except Exception:
                sys.path.remove(c)
            break
    return _REAL_REF or None


def _cpu_worker(job):
    m, q, v, eps, rr, rv = job
    ref = _real_reference()
    if ref is not None:
        # the reference's own path: batch_stability_analyzer.py:62-80 -> stability_analyzer.py:69-259
        import contextlib
        import io
        from minbody.simulation import NBodySimulation
        from minbody.batch_stability_analyzer import BatchStabilityAnalyzer
        with contextlib.redirect_stdout(io.StringIO()):
            sim = NBodySimulation(masses=list(m), positions=[tuple(x) for x in q], velocities=[tuple(x) for x in v],
                                  softening=float(eps), integrator_mode=MODE)
            BatchStabilityAnalyzer(n_steps=N_STEPS, dt=DT, mode="full").analyze_batch([sim], show_progress=False)
        return STEPS_PER_SYSTEM
    from oracle import nbody_oracle as O
    sim = O.OracleSim(m, q, v, softening=float(eps), integrator_mode=MODE)
    O.run_stability_analysis(sim, N_STEPS, DT, "full", rr, rv)
    return STEPS_PER_SYSTEM


def cpu_kind():
    return "reference" if _real_reference() is not None else "port"


def cpu_sample_jobs(n_jobs: int, seed: int):
    """A bounded sample WITH THE WORKLOAD'S OWN MIX: the first n_jobs systems of the same diverse generator (46 % of
    them N = 3: hierarchical triples + a share of the random / polygon / close-encounter cohorts), shuffled so that
    every worker sees the same mix."""
    inp = make_inputs(n_jobs, seed)
    jobs = []
    for N in sorted(inp):
        d = inp[N]
        for i in range(d["m"].shape[0]):
            jobs.append((d["m"][i], d["q"][i], d["v"][i], d["eps"][i], d["raw_dr"][i], d["raw_dv"][i]))
    np.random.default_rng(seed).shuffle(jobs)
    return jobs[:n_jobs]


class CpuPool:
    """Worker processes created ONCE, outside every timed interval."""

    def __init__(self, cores: int):
        import multiprocessing as mp
        restore_affinity()
        self.cores = cores
        self.pool = mp.get_context("fork").Pool(cores) if cores > 1 else None
        if self.pool is not None:                      # start-up cost (fork + imports) is paid here
            self.pool.map(_warm_worker, range(cores))

    def run(self, fn, jobs):
        t0 = time.perf_counter()
        if self.pool is not None:
            done = sum(self.pool.map(fn, jobs, chunksize=max(1, len(jobs) // (4 * self.cores))))
        else:
            done = sum(fn(j) for j in jobs)
        dt = time.perf_counter() - t0
        return done / dt, dt

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def _warm_worker(_):
    _real_reference()
    from oracle import nbody_oracle, hamsoft_oracle  # noqa: F401
    return 0


def cpu_pairs_per_s(n: int = 4096, reps: int = 3):
    """CPU baseline for the large-N metric: dense gravitational_force (forces.py:63-75; the reference's own function
    when a checkout is importable, else the oracle restatement) at the largest N whose (N,N,2) fp64 temporaries fit
    comfortably; pairs/s on one core (the reference is single-threaded)."""
    from nbodysimproject_b200.largen import make_disc
    m, q, _ = make_disc(n, seed=5)
    ref = _real_reference()
    if ref is not None:
        from minbody.forces import gravitational_force
        fn = lambda: gravitational_force(q, m, eps=1e-3, G=1.0)
    else:
        from oracle import nbody_oracle as O
        fn = lambda: O.accelerations(q, m, 1e-3, 1.0)
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
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
        "cpu_baseline": {"value": value, "unit": "pair-interactions/s", "cores": 1, "kind": cpu_kind(),
                         "sample": "dense gravitational_force at N=4096, 3 calls per step"},
        "e2e": {"value": value, "unit": "pair-interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def impl_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "largen":
        return impl_reference_largen(args)
    cores = os.cpu_count() or 1
    n_jobs = cores * 48                       # 48 systems per core and step: the per-process tail stays small
    pool = CpuPool(cores)
    times, done = [], 0
    for it in range(args.warmup + args.steps):
        jobs = cpu_sample_jobs(n_jobs, seed=1000 + it)
        rate, dt = pool.run(_cpu_worker, jobs)
        if it >= args.warmup:
            times.append(dt)
            done += n_jobs * STEPS_PER_SYSTEM
        if sum(times) > 150:          # keep the whole arm within a few minutes
            break
    pool.close()
    total = sum(times)
    value = done / total
    k = len(times)
    kind = cpu_kind()
    line = {
        "impl": "reference", "metric": "system-steps/s (N=3-8 ensembles, MEGNO on)", "value": value,
        "unit": "system-steps/s", "n_gpus": args.gpus, "steps": k, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(k, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 batch stability ensemble: diverse N=3-8 cohort, yoshida4, 1000+50 MEGNO steps, "
                               "mode full (bounded CPU sample of the same generator, workload-weighted N mix)",
                   "systems_per_step": n_jobs, "integrator": MODE, "dt": DT},
        "cpu_baseline": {"value": value, "unit": "system-steps/s", "cores": cores, "kind": kind,
                         "sample": f"{n_jobs} systems x {STEPS_PER_SYSTEM} steps per bench step, "
                                   + ("the reference's BatchStabilityAnalyzer.analyze_batch" if kind == "reference"
                                      else "NumPy oracle (oracle/nbody_oracle.py; measured 3.0x faster per core than the reference's "
                                           "own analyze_batch in the build container, 22.1e3 vs 7.4e3 system-steps/s on 8 cores, "
                                           "DESIGN.md section 5)")
                                   + f" in {cores} worker processes created before the timed steps"},
        "e2e": {"value": value, "unit": "system-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------

NOMINAL_FP64 = 64 * 2 * 148 * 1.965e9 * 1e-12      # 64 DFMA/clk/SM x 148 SMs x 1.965 GHz = 37.2 TFLOP/s
NOMINAL_FP32 = 128 * 2 * 148 * 1.965e9 * 1e-12     # 128 FFMA/clk/SM                      = 74.5 TFLOP/s


def _traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this command
    (profiles/r2_traffic.json, written by tools/ncu_traffic.py); None when no capture is on file."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            t = json.load(f)
        return t.get(key)
    except Exception:
        return None


def _barrier(torch, dist, world):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def _allmax(torch, dist, world, dev, *vals):
    if world == 1:
        return vals
    tt = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return tuple(float(x) for x in tt)


def impl_b200(args):
    import torch
    import torch.distributed as dist
    from nbodysimproject_b200 import _lib as L

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
    L.load()
    ctx = dict(torch=torch, dist=dist, world=world, rank=rank, local=local, dev=dev, numa=numa)

    if args.workload == "largen":
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        line = largen_section(args, ctx, steps=args.steps, sampler=sampler, with_hamsoft=True, standalone=True)
    elif args.workload == "c2":
        line = bench_c2(args, torch, dist, world, rank, local, dev)
    elif args.workload in ("c4", "c1"):
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        line = secondary_section(args, ctx, args.workload, steps=args.steps, sampler=sampler, standalone=True)
    else:
        line = ensemble_section(args, ctx)
    if rank == 0 and line is not None:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ensemble_section(args, ctx):
    """Default line: C3 (the configuration BASELINE.json's metric is quoted on) + compact sections for the second half
    of the metric (large N) and the other configs (C1 ham_soft, C4 whfast), + the cross-rank consistency checks."""
    torch, dist, world, rank, local, dev = (ctx[k] for k in ("torch", "dist", "world", "rank", "local", "dev"))
    from nbodysimproject_b200 import _lib as L
    from nbodysimproject_b200 import ensemble as E
    lib = L.load()
    B_total = args.systems
    inp = make_inputs(B_total, seed=42 + rank)
    Ns = sorted(inp, reverse=True)   # launch the large-N buckets first: their sequential sub-step tails are the longest
    host, devb, devb2 = {}, {}, {}
    h2d = 0
    for N in Ns:
        d = inp[N]
        B = d["m"].shape[0]
        hb = {}
        for k in ("m", "q", "v", "eps", "raw_dr", "raw_dv"):
            t = torch.from_numpy(np.ascontiguousarray(d[k], dtype=np.float64)).pin_memory()
            hb[k] = t
            h2d += t.numel() * 8
        hb["stat"] = torch.empty((B, L.N_STATIC), dtype=torch.float64).pin_memory()
        hb["nsub"] = torch.empty((B,), dtype=torch.int32).pin_memory()
        hb["status"] = torch.empty((B,), dtype=torch.int32).pin_memory()
        host[N] = hb
        sets = []
        for _ in range(max(1, args.sets)):   # buffer / stream sets: consecutive steps alternate between them
            bk = E.DeviceBucket(hb["m"], hb["q"], hb["v"], hb["eps"], 1.0, MODE, dev)
            if sets:
                bk.v0, bk.q0, bk.rdr, bk.rdv = sets[0].v0, sets[0].q0, sets[0].rdr, sets[0].rdv
            else:
                bk.v0 = bk.v.clone()
                bk.q0 = bk.q.clone()
                bk.rdr = hb["raw_dr"].to(dev)
                bk.rdv = hb["raw_dv"].to(dev)
            # highest priority, like the slot streams of the host entry point: nb_ensemble_run_f64 then demotes only the
            # BULK of each main kernel to an internal normal-priority stream, so every latency-bound piece (the
            # sub-step-heavy heads, the MEGNO kernels with their own tails, energy / finalize) is dispatched first
            bk.stream = torch.cuda.Stream(device=dev, priority=-1)
            sets.append(bk)
        devb2[N] = sets
        devb[N] = sets[0]
    del inp
    prep_flags = L.PREP_REMOVE_COM | L.PREP_CTOR_KICK | L.PREP_SNAPSHOT_KICK
    interval = max(1, N_STEPS // 100)
    launches_per_step = 0
    # per timed step and bucket: {earliest CTA start, latest warp end} of the main-phase kernels in %globaltimer ns,
    # published by the kernels themselves (nb_ensemble_run_counted_f64, t_main)
    I64MAX = (1 << 63) - 1
    stamps = {N: torch.tensor([[I64MAX, 0]] * (args.steps + 1), dtype=torch.int64, device=dev) for N in Ns}
    timed_step = [0]

    step_no = [0]
    inflight = []
    # launch order of the buckets: the large-N buckets first (their sequential sub-step tails are the longest)
    issue_order = [int(x) for x in args.order.split(",")] if args.order else list(Ns)

    def step_device(record=False, first=False, last=False):
        """One step = one full pass of the analysis over every bucket.  Consecutive steps are independent batches and are
        enqueued back to back on the bucket streams WITHOUT a device-wide join in between (each bucket's stream keeps its
        own steps in order), so the sub-step-heavy tail of one bucket overlaps the next step of the others, as in any
        pipelined deployment; the timed region starts with a fork from the timing stream and ends with a join into it.
        (Alternating two buffer / stream sets, as the e2e leg does through the host entry point's slots, was tried and
        made the device-resident step time launch-order sensitive, 54-90 ms; one set is stable at ~55 ms.)"""
        nonlocal launches_per_step
        cur = torch.cuda.current_stream()
        n = 0
        par = step_no[0] % max(1, args.sets)
        step_no[0] += 1
        # the construction-time kernels of every bucket first (0.3 % of the step), then the runs: nb_ensemble_run_f64
        # launches each bucket's sub-step-heavy head at high priority, so no head waits behind another bucket's bulk
        for N in issue_order:
            bk = devb2[N][par]
            if first:
                for b2 in devb2[N]:
                    b2.stream.wait_stream(cur)
            with torch.cuda.stream(bk.stream):
                bk.q.copy_(bk.q0)
                bk.v.copy_(bk.v0)
                bk.prepare(prep_flags, 0.01, 0.01, DT, 50, want_static=True)     # 1 kernel
                bk.sort(args.heavy_threshold)                                    # 3 kernels
        for N in issue_order:
            bk = devb2[N][par]
            with torch.cuda.stream(bk.stream):
                ts = stamps[N][timed_step[0]] if record else None
                bk.dyn = bk.run(DT, N_STEPS, interval, N_MEGNO, bk.rdr, bk.rdv, flags=L.RUN_ENERGY, t_main=ts)  # 2+2+1+1 kernels
                n += 10
            devb[N] = bk                  # the bucket of the most recent step (cross-checks below)
        if args.sets > 1:
            # a streaming caller keeps a bounded number of steps in flight: step k is enqueued, then the host waits for
            # step k - (sets - 1) (event waits on the bucket streams, no device-wide join) -- exactly what the e2e leg does
            # through nb_host_sync
            evs = []
            for N in issue_order:
                ev = torch.cuda.Event()
                ev.record(devb2[N][par].stream)
                evs.append(ev)
            inflight.append(evs)
            while len(inflight) > args.sets - 1:
                for ev in inflight.pop(0):
                    ev.synchronize()
        if last:
            for N in Ns:
                for b2 in devb2[N]:
                    cur.wait_stream(b2.stream)
        launches_per_step = n
        if record:
            timed_step[0] += 1

    # ---- e2e: the host entry point as a streaming caller uses it.  Every step has its own host buffers for everything
    # the call writes (the kicked velocities go back into the caller's v like the reference's snapshot() mutation, plus the
    # two feature tables), and TWO steps are in flight on two sets of workspace slots: step k+1's inputs and step k-1's
    # results move while step k computes.  Every step's inputs cross PCIe inside the timed region.
    import ctypes
    n_e2e = max(1, min(args.warmup, 2)) + args.steps
    e2e_opts = L.HostOpts(flags=L.HOST_COMPACT_DYN)
    order = sorted(Ns, key=lambda n: host[n]["m"].numel())      # smallest H2D first: the GPU starts computing early
    for N in Ns:
        hb = host[N]
        B = hb["m"].shape[0]
        hb["v_step"] = [hb["v"].clone().pin_memory() for _ in range(n_e2e)]
        hb["dyn_step"] = [torch.empty((B, L.N_DYN_USER), dtype=torch.float64).pin_memory() for _ in range(2)]
        hb["stat_step"] = [hb["stat"], torch.empty_like(hb["stat"]).pin_memory()]
        hb["nsub_step"] = [hb["nsub"], torch.empty_like(hb["nsub"]).pin_memory()]
        hb["status_step"] = [hb["status"], torch.empty_like(hb["status"]).pin_memory()]
    d2h_c = sum(host[N]["dyn_step"][0].numel() * 8 + host[N]["stat"].numel() * 8 + host[N]["m"].shape[0] * 8
                + host[N]["v"].numel() * 8 for N in Ns)

    def issue_e2e(k):
        par = k & 1
        for j, N in enumerate(order):
            hb = host[N]
            B = hb["m"].shape[0]
            L.check(lib.nb_ensemble_analyze_host_ex(
                L.ptr(hb["m"]), L.ptr(hb["q"]), L.ptr(hb["v_step"][k]), L.ptr(hb["eps"]), 1.0, B, N, L.MODES[MODE],
                prep_flags, 0.01, 0.01, DT, N_STEPS, N_MEGNO, 50, L.ptr(hb["raw_dr"]), L.ptr(hb["raw_dv"]),
                L.ptr(hb["dyn_step"][par]), L.ptr(hb["stat_step"][par]), L.ptr(hb["nsub_step"][par]),
                L.ptr(hb["status_step"][par]), local, 8 * par + j, ctypes.byref(e2e_opts)),
                "nb_ensemble_analyze_host_ex")

    def sync_e2e(k):
        for j in range(len(order)):
            L.check(lib.nb_host_sync(8 * (k & 1) + j), "nb_host_sync")

    def run_e2e(k0, k1):
        """steps k0 .. k1-1, two in flight; returns the wall time of exactly that"""
        t0 = time.perf_counter()
        for k in range(k0, k1):
            issue_e2e(k)
            if k > k0:
                sync_e2e(k - 1)
        sync_e2e(k1 - 1)
        return time.perf_counter() - t0

    # ---- value: device-resident
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for k in range(args.warmup):
        step_device(first=(k == 0), last=(k == args.warmup - 1))
    _barrier(torch, dist, world)
    mark0 = sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        step_device(record=True, first=(k == 0), last=(k == args.steps - 1))
    e1.record()
    _barrier(torch, dist, world)
    mark1 = sampler.mark()
    t_dev = e0.elapsed_time(e1) * 1e-3
    # ---- e2e: host buffers through the C ABI
    n_warm = n_e2e - args.steps
    run_e2e(0, n_warm)
    _barrier(torch, dist, world)
    t_e2e = run_e2e(n_warm, n_e2e)
    torch.cuda.synchronize()
    _barrier(torch, dist, world)
    clocks = sampler.stop(mark0, mark1) if rank == 0 else None
    t_dev, t_e2e = _allmax(torch, dist, world, dev, t_dev, t_e2e)
    sys_steps = float(B_total) * world * STEPS_PER_SYSTEM * args.steps

    # ---- cross-check: e2e and device paths give identical feature tables; count statuses
    last = (n_e2e - 1) & 1
    same = all(np.array_equal(host[N]["dyn_step"][last].numpy(), devb[N].dyn[:, :L.N_DYN_USER].cpu().numpy(), equal_nan=True)
               and np.array_equal(host[N]["stat_step"][last].numpy(), devb[N].static.cpu().numpy(), equal_nan=True)
               for N in Ns)
    n_bad = int(sum(int((host[N]["status_step"][last].numpy() != 0).sum()) for N in Ns))

    # ---- roofline of the dominant kernel family: ensemble_main_kernel<N, yoshida4>, N = 3..8, timed INSIDE the timed
    # steps: the C ABI records a CUDA-event pair around each bucket's main-phase launches (head + rest); the six
    # buckets run concurrently on their streams, so the figure per step is the algorithmic flops of the six launches
    # over the window [first begin, last end].
    roof = None
    if rank == 0:
        peak = L.peak_flops(0, local)
        tot_fl, meg_fl, per = 0.0, 0.0, []
        for N in Ns:
            bk = devb[N]
            nsub_sum = int(bk.n_sub.sum().item())       # n_sub of the last timed step (deterministic in the inputs)
            bk.flops = flops_main(N, nsub_sum, N_STEPS)
            tot_fl += bk.flops
            meg_fl += flops_megno(N, bk.B, nsub_sum, N_MEGNO)
            ts = stamps[N][:args.steps].cpu().numpy()
            bk.ts = ts
            per.append(dict(N=N, B=bk.B, mean_n_sub=nsub_sum / bk.B, max_n_sub=int(bk.n_sub.max().item()),
                            in_step_ms=float(np.mean(ts[:, 1] - ts[:, 0])) * 1e-6, flops=bk.flops))
        # steps are pipelined, so their windows overlap: the main-phase time of the timed region is the span from the
        # earliest start of the first step to the latest end of the last one, divided by the number of steps
        span = max(devb[N].ts[:, 1].max() for N in Ns) - min(devb[N].ts[:, 0].min() for N in Ns)
        win = float(span) * 1e-9 / args.steps
        # the span also contains the MEGNO kernels (and prepare / sort / energy / finalize) of every timed step but the
        # last, whose MEGNO phase starts after the last main-phase stamp: their algorithmic flops belong to the same
        # window (SURVEY.md 8d states the C3 unit as "system-step, yoshida4 + MEGNO")
        meg_in = meg_fl * (args.steps - 1) / float(args.steps)
        ach_main = tot_fl / win * 1e-12
        ach = (tot_fl + meg_in) / win * 1e-12
        roof = {"bound": "fp64", "kernel": "ensemble_main_kernel + ensemble_megno_kernel <N=3..8, yoshida4> "
                                           "(6 + 6 concurrent launches per step)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "frac_nominal": ach / NOMINAL_FP64, "peak_nominal": NOMINAL_FP64,
                "main_phase_only": {"achieved": ach_main, "frac": ach_main / peak, "flops_per_step": tot_fl,
                                    "note": "r1 / early-r2 accounting: main-phase flops over the same window (which "
                                            "also holds the MEGNO kernels' time)"},
                "flops_megno_phase_per_step": meg_fl,
                "traffic": (lambda t: None if t is None else t * (B_total / float(1 << 20)))(
                    _traffic("ensemble_main_six_launches_bytes_per_2p20_systems")),
                "traffic_note": "dram__bytes_read + dram__bytes_write over the main-phase launches of one step from the "
                                "committed ncu capture of this command at 2^20 systems/GPU (profiles/r2_traffic.json, "
                                "tools/ncu_traffic.py), scaled by the batch size; the state is read once and lives in "
                                "registers (algorithmic bytes: ~0.45 KB in + 0.18 KB out per system = 0.66 GB/step)",
                "flops_per_step": tot_fl + meg_in, "ms": win * 1e3,
                "timing": "%globaltimer stamps published by the main kernels themselves (first CTA start, last warp end; "
                          "nb_ensemble_run_counted_f64 t_main) inside the timed steps; span from the earliest start of the "
                          "first timed step to the latest end of the last one over all six buckets, divided by the steps "
                          "(consecutive steps are pipelined, so per-step windows overlap)",
                "peak_source": "nb_peak_flops(0): register-resident DFMA micro-benchmark, same GPU, same run "
                               "(MEASURED_PEAKS.json has no FP64 figure; nominal 64 DFMA/clk/SM x 148 x 1.965 GHz = 37.2)",
                "flop_model": "SURVEY.md 8d: per sub-step 3 x 14 N(N-1) + 36 N; per MEGNO step additionally 19 N(N-1) + 8 N "
                              "for the tangent map",
                "share_of_step": win / (t_dev / args.steps), "per_bucket": per}

    # ---- N > 1: cross-rank correctness (the reference's contract: per-system results independent of batching,
    # batch_stability_analyzer.py:62-80)
    checks = {"e2e_equals_device_path": bool(same), "systems_with_nonzero_status": n_bad}
    if world > 1:
        checks.update(consistency_checks(ctx))

    # free the C3 buffers before the secondary sections
    for N in Ns:
        devb[N] = None
        devb2[N] = None
        host[N] = None
    torch.cuda.empty_cache()

    largen = c1 = c4 = None
    if not args.no_largen:
        largen = largen_section(args, ctx, steps=3, sampler=None, with_hamsoft=not args.no_largen_hamsoft, standalone=False)
    if not args.no_secondary:
        c1 = secondary_section(args, ctx, "c1", steps=3, sampler=None, standalone=False)
        c4 = secondary_section(args, ctx, "c4", steps=3, sampler=None, standalone=False)

    # ---- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        n_jobs = cores * 96                  # ~10-15 s of CPU work on the box's cores
        pool = CpuPool(cores)
        rate, dt_cpu = pool.run(_cpu_worker, cpu_sample_jobs(n_jobs, 777))
        pool.close()
        kind = cpu_kind()
        cpu = {"value": rate, "unit": "system-steps/s", "cores": cores, "kind": kind,
               "sample": f"{n_jobs} systems x {STEPS_PER_SYSTEM} steps of the same generator (workload-weighted N mix), "
                         + ("the reference's BatchStabilityAnalyzer" if kind == "reference" else
                            "NumPy oracle (3.0x faster per core than the reference itself, measured in the build container)")
                         + f" in {cores} processes, {dt_cpu:.1f} s"}
    if rank != 0:
        return None
    return {
        "metric": "system-steps/s (N=3-8 ensembles, MEGNO on)", "value": sys_steps / t_dev,
        "unit": "system-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 batch stability ensemble (BASELINE.json configs[2]): diverse cohort "
                               "40% random N=3-8 / 30% hierarchical triples / 20% polygons / 10% close encounters, "
                               "yoshida4 dt=0.01, 1000 steps + 50 tangent-map MEGNO steps, mode full",
                   "systems_per_gpu": B_total, "buckets": {str(p["N"]): int(p["B"]) for p in (roof or {}).get("per_bucket", [])},
                   "sharding": "by system, no collective", "cpu_cores_bound_to_gpu_numa_node": ctx["numa"],
                   "step_pipelining": "steps are independent batches enqueued back to back on the bucket streams, no "
                                      "device-wide join between steps (each bucket stream keeps its own steps in order)",
                   "l2_note": f"inputs re-read from HBM each step ({h2d / 1e6:.0f} MB per GPU > 126 MB L2)"},
        "e2e": {"value": sys_steps / t_e2e, "unit": "system-steps/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h_c), "ms_per_step": 1e3 * t_e2e / args.steps,
                "api": "nb_ensemble_analyze_host_ex (C ABI, pinned host buffers; returns the kicked velocities, the 17 "
                       "user-visible dynamic columns, the 25 static columns, n_sub and status per system)",
                "pipelining": "two steps in flight on two sets of workspace slots (fresh host buffers per step): a step's "
                              "H2D / D2H overlap the neighbouring steps' kernels; wall clock over all timed steps"},
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "largen": largen, "c1": c1, "c4": c4,
        "checks": checks,
    }


def consistency_checks(ctx):
    """world > 1, every rank: (a) the SAME 4,096-system ensemble analysed in full on every rank -> the feature tables must
    be bit-identical across ranks; (b) the same ensemble sharded by system over the ranks and gathered -> must equal
    rank 0's full table bit for bit (results independent of batching / GPU count); (c) ham_soft: eps, pi of the same
    256 systems equal across ranks."""
    torch, dist, world, rank, dev = (ctx[k] for k in ("torch", "dist", "world", "rank", "dev"))
    from nbodysimproject_b200 import ensemble as E
    from nbodysimproject_b200 import hamsoft as H
    from nbodysimproject_b200 import sharding as S
    inp = make_inputs(4096, seed=20261)
    ident, sharded = True, True
    for N in sorted(inp):
        d = inp[N]
        B = d["m"].shape[0]

        def compute(lo, hi):
            r = E.analyze_bucket(d["m"][lo:hi], d["q"][lo:hi], d["v"][lo:hi], d["eps"][lo:hi], 1.0, MODE, 200, DT,
                                 "full", d["raw_dr"][lo:hi], d["raw_dv"][lo:hi], device=dev)
            return np.concatenate([r.dyn, r.static], axis=1)

        full = compute(0, B)
        t = torch.from_numpy(full.view(np.int64).copy()).to(dev)
        lo_t, hi_t = t.clone(), t.clone()
        dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
        ident &= bool(torch.equal(lo_t, hi_t))
        gathered = S.analyze_sharded(compute, B)
        if rank == 0:
            sharded &= bool(np.array_equal(gathered, full, equal_nan=True))
    # ham_soft
    m, q, v = _c1_inputs(256, 5)
    hs, s0 = H.default_params(object(), 1e-3, 1e-4, 256)
    hb = H.HamSoftBucket(m, q, v, hs, np.stack([np.maximum(s0, hs[:, H.P["eps_min"]]), np.zeros(256)], 1), 1.0, dev)
    hb.setup(calibrate=True, freeze_dt=0.01)
    hb.run(0.01, 50)
    t = hb.eps_pi.clone().view(torch.int64)
    lo_t, hi_t = t.clone(), t.clone()
    dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
    hs_ident = bool(torch.equal(lo_t, hi_t))
    flag = torch.tensor([1 if sharded else 0], device=dev)
    dist.broadcast(flag, 0)
    return {"ranks_bit_identical": bool(ident), "sharded_equals_single_bit_identical": bool(int(flag[0]) == 1),
            "hamsoft_eps_pi_identical_across_ranks": hs_ident,
            "consistency_sample": "4096 systems (same generator, fixed seed), yoshida4, 200 + MEGNO steps, full feature tables"}


# ---------------------------------------------------------------------------------------------
# large-N direct sum (second half of BASELINE.json's metric)
# ---------------------------------------------------------------------------------------------

def largen_section(args, ctx, steps, sampler, with_hamsoft, standalone):
    """pair-interactions/s of one force evaluation over all ordered pairs of an N-particle system (strong scaling: N
    fixed, i-blocks sharded, one in-place position all-gather per evaluation); e2e with host positions in / host
    accelerations out; fp64 spot check of 48 random particles; wall time of one full ham_soft Strang sub-step."""
    torch, dist, world, rank, local, dev = (ctx[k] for k in ("torch", "dist", "world", "rank", "local", "dev"))
    from nbodysimproject_b200 import _lib as L
    from nbodysimproject_b200 import largen as LN
    n = int(args.n)
    m, q, v = LN.make_disc(n, seed=1)
    sim = LN.LargeNSimulation(m, q, v, G=1.0, softening=1e-3, device=dev)
    mark0 = sampler.mark() if sampler is not None else 0
    t = LN.measure_force(sim, steps, args.warmup)
    mark1 = sampler.mark() if sampler is not None else 0
    # e2e: host positions in, host accelerations out, every step
    xym_h = sim.xym.cpu().pin_memory()
    acc_h = torch.empty((sim.ni, 2), dtype=torch.float32).pin_memory()
    _barrier(torch, dist, world)
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.xym.copy_(xym_h, non_blocking=True)
        sim.accelerations()
        acc_h.copy_(sim.acc, non_blocking=True)
        torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    (t_e2e,) = _allmax(torch, dist, world, dev, t_e2e)
    # fp64 direct sum for 48 random particles of THIS rank's i-block (forces.py:63-75 in fp64 on the device)
    g = torch.Generator(device="cpu").manual_seed(7 + rank)
    idx = (torch.randperm(sim.ni, generator=g)[:48] + sim.i0).to(dev)
    x64 = sim.xym[:, 0:2].double()
    m64 = sim.xym[:, 2].double()
    d = x64[idx][:, None, :] - x64[None, :, :]
    w = (d * d).sum(-1) + 1e-3 ** 2
    w[torch.arange(48, device=dev), idx] = float("inf")
    ref = -(m64[None, :, None] * d * (w ** -1.5)[:, :, None]).sum(1)
    got = sim.acc[idx - sim.i0].double()
    err = float(((got - ref).norm(dim=1) / ref.norm(dim=1)).max())
    (err,) = _allmax(torch, dist, world, dev, err)
    del d, w, x64
    strang = None
    if with_hamsoft:
        n_hs = int(getattr(args, "n_hamsoft", 0) or n)            # C5 as BASELINE.json words it: the same N = 2^20 particles
        mh, qh, vh = LN.make_disc(n_hs, seed=1)
        hs = LN.LargeNHamSoftSimulation(mh, qh, vh, softening=2.0 / math.sqrt(n_hs), initial_dt=1e-3, device=dev)
        hsub = 1e-3 / hs.frozen_n_sub
        hs.strang_step(hsub)
        _barrier(torch, dist, world)
        p0, f0, u0 = hs.n_passes, hs.force_evals, getattr(hs, "n_full_passes", 0)
        t0 = time.perf_counter()
        hs.strang_step(hsub)
        torch.cuda.synchronize()
        ts = time.perf_counter() - t0
        (ts,) = _allmax(torch, dist, world, dev, ts)
        n_pass = (hs.n_passes - p0) + (hs.force_evals - f0)
        n_full = (getattr(hs, "n_full_passes", 0) - u0) + (hs.force_evals - f0)
        # ranks must agree on the replicated scalars
        sc = torch.tensor([hs.eps, hs.pi], dtype=torch.float64, device=dev).view(torch.int64)
        same = True
        if world > 1:
            lo_t, hi_t = sc.clone(), sc.clone()
            dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
            same = bool(torch.equal(lo_t, hi_t))
        strang = {"n": n_hs, "ms": 1e3 * ts, "pair_passes": n_pass, "full_n2_passes": n_full,
                  "locality_culled_passes": n_pass - n_full, "solver_sweeps": hs.last_sweeps,
                  "note": "S V T V S of the adaptive-epsilon flow: 2 force evaluations + 1 long-range legacy-direction pass "
                          "over all N^2 pairs; the eps* solver sweeps / density / gradient passes skip tiles beyond 9.35 h "
                          "(exact zeros of ex2.approx.ftz) on the Morton-ordered particle array",
                  "algorithmic_pair_evaluations_per_s": n_pass * float(n_hs) * n_hs / ts, "frozen_n_sub": hs.frozen_n_sub,
                  "eps": hs.eps, "pi": hs.pi, "eps_min": hs.eps_min, "eps_max": hs.eps_max,
                  "eps_pi_identical_across_ranks": same}
        del hs
    if rank != 0:
        return None
    pairs = float(n) * float(n) * steps
    peak32 = L.peak_flops(1, local)
    ach = 14.0 * pairs / t * 1e-12 / world
    cpu = None
    if world == 1 and not args.no_cpu:
        restore_affinity()
        rate, dtc = cpu_pairs_per_s(4096, 5)
        cpu = {"value": rate, "unit": "pair-interactions/s", "cores": 1, "kind": cpu_kind(),
               "sample": f"dense gravitational_force at N=4096 ({dtc * 1e3:.0f} ms per call; the (N,N,2) fp64 temporaries "
                         "make N=2^20 impossible on the CPU path: 17.6 TB)"}
    line = {
        "metric": "pair-interactions/s at N=2^20", "value": pairs / t, "unit": "pair-interactions/s",
        "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C5 single large-N direct sum, N={n}, Plummer softening 1e-3, one force evaluation "
                               "(+ in-place NCCL position all-gather when sharded) per step",
                   "l2_note": "j-array (16 B x N) is L2-resident by design; inputs are 16 MB at N=2^20"},
        "e2e": {"value": pairs / t_e2e, "unit": "pair-interactions/s", "h2d_bytes_per_step": int(n * 16),
                "d2h_bytes_per_step": int(sim.ni * 8), "ms_per_step": 1e3 * t_e2e / steps},
        "gpu_launches": 2 * steps,
        "roofline": {"bound": "fp32", "kernel": "largeN_accel_x2_kernel", "achieved": ach, "peak": peak32,
                     "unit": "TFLOP/s", "frac": ach / peak32, "frac_nominal": ach / NOMINAL_FP32,
                     "peak_nominal": NOMINAL_FP32, "traffic": _traffic("largeN_accel_bytes_per_launch_2p20"),
                     "flops_per_pair": 14,
                     "peak_source": "nb_peak_flops(1): register-resident FFMA micro-benchmark, same GPU, same run"},
        "checks": {"max_rel_err_vs_fp64_direct_sum": err, "sample": "48 random particles per rank, max over ranks"},
        "hamsoft_strang_substep": strang, "cpu_baseline": cpu,
    }
    if standalone:
        line["clocks"] = sampler.stop(mark0, mark1) if sampler is not None else None
    return line


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
    m, q, v, n_steps, dt = job
    if _real_reference() is not None:
        from minbody.simulation import NBodySimulation
        sim = NBodySimulation(masses=list(m), positions=[tuple(x) for x in q], velocities=[tuple(x) for x in v],
                              softening=0.0, integrator_mode="whfast")
    else:
        from oracle import nbody_oracle as O
        sim = O.OracleSim(m, q, v, softening=0.0, integrator_mode="whfast")
    for _ in range(n_steps):
        sim.step(dt)
    return n_steps


def _cpu_c1(job):
    m, q, v, n_steps, dt = job
    if _real_reference() is not None:
        from minbody.simulation import NBodySimulation
        sim = NBodySimulation(masses=list(m), positions=[tuple(x) for x in q], velocities=[tuple(x) for x in v],
                              softening=1e-3, integrator_mode="ham_soft")
    else:
        from oracle.hamsoft_oracle import HamSoftOracleSim
        sim = HamSoftOracleSim(m, q, v, softening=1e-3)
    for _ in range(n_steps):
        sim.step(dt)
    return n_steps


def _c1_api_single_system(n_steps=1000):
    """BASELINE.json configs[0] exactly as worded: ONE NBodySimulation, `for _ in range(1000): sim.step(0.01)`
    (simulation.py:667-676).  A single 3-body system cannot fill a GPU: every step() is one H2D + one kernel launch
    of ONE warp + one D2H, so this number is launch-latency, reported beside the batched figure, not instead of it."""
    import torch
    from nbodysimproject_b200 import NBodySimulation

    def make():
        return NBodySimulation(masses=[1.0, 0.5, 0.1], positions=[(0.0, 0.0), (1.0, 0.0), (2.0, 0.0)],
                               velocities=[(0.0, 0.0), (0.0, 1.0), (0.0, 0.5)], integrator_mode="ham_soft")
    sim = make()
    for _ in range(20):
        sim.step(0.01)
    sim = make()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_steps):
        sim.step(0.01)
    torch.cuda.synchronize()
    t_loop = time.perf_counter() - t0
    q_loop = sim.pos.copy()
    sim2 = make()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim2.step_many(0.01, n_steps)
    torch.cuda.synchronize()
    t_many = time.perf_counter() - t0
    return {"steps": n_steps, "step_loop_steps_per_s": n_steps / t_loop, "step_loop_s": t_loop,
            "step_many_steps_per_s": n_steps / t_many, "step_many_s": t_many,
            "step_many_equals_step_loop": bool(np.array_equal(q_loop, sim2.pos)),
            "reference_cpu_steps_per_s_BASELINE_md": 189.0,
            "note": "one NBodySimulation(...); for _ in range(1000): sim.step(0.01) -- launch-latency bound (one warp)"}


def secondary_section(args, ctx, which, steps, sampler, standalone):
    """c4 | c1: system-steps/s of the whfast planetary ensemble / the batched README ham_soft system.
    One bench step = n_steps integrator steps of every system in one persistent-kernel launch per bucket.
    The roofline uses COUNTED work returned by the kernels (Newton iterations / Jacobi sweeps), SURVEY.md 8d."""
    torch, dist, world, rank, local, dev = (ctx[k] for k in ("torch", "dist", "world", "rank", "local", "dev"))
    from nbodysimproject_b200 import _lib as L
    from nbodysimproject_b200 import ensemble as E
    from nbodysimproject_b200 import hamsoft as H
    c4 = which == "c4"
    n_steps = int(args.horizon) if (standalone and args.horizon) else 1000
    dt = 0.01 * 2.0 * np.pi if c4 else 0.01
    B = args.systems if c4 else min(args.systems, 1 << 17)
    if standalone and args.horizon and args.horizon > 1000 and args.systems == (1 << 20):
        B = 1 << 16                                   # long horizons: 65,536 systems (SURVEY.md 8d C4 / VERDICT r1)
    runs = []
    if c4:
        inp = _c4_inputs(B, 42 + rank)
        for N, (m, q, v, eps) in sorted(inp.items()):
            bk = E.DeviceBucket(m, q, v, eps, 1.0, "whfast", dev)
            bk.q0, bk.v0 = bk.q.clone(), bk.v.clone()
            bk.stream = torch.cuda.Stream(device=dev)
            bk.work = torch.zeros((bk.B, 2), dtype=torch.float64, device=dev)
            runs.append(bk)
    else:
        m, q, v = _c1_inputs(B, 42 + rank)
        hs, s0 = H.default_params(object(), 1e-3, 1e-4, B)
        ep = np.stack([np.maximum(s0, hs[:, H.P["eps_min"]]), np.zeros(B)], 1)
        hb = H.HamSoftBucket(m, q, v, hs, ep, 1.0, dev)
        hb.setup(calibrate=True, freeze_dt=dt)
        hb.sort()
        hb.q0, hb.v0, hb.ep0, hb.hs0 = hb.bk.q.clone(), hb.bk.v.clone(), hb.eps_pi.clone(), hb.hs.clone()
        hb.work = torch.zeros((B, 2), dtype=torch.float64, device=dev)
        runs.append(hb)
    I64MAX = (1 << 63) - 1
    for r in runs:
        r.stamps = torch.tensor([[I64MAX, 0]] * (steps + 1), dtype=torch.int64, device=dev)
    timed_step = [0]

    def step(record=False):
        cur = torch.cuda.current_stream()
        if c4:
            for bk in runs:
                bk.stream.wait_stream(cur)
                with torch.cuda.stream(bk.stream):
                    bk.q.copy_(bk.q0); bk.v.copy_(bk.v0)
                    bk.prepare(L.PREP_REMOVE_COM | L.PREP_CTOR_KICK, dt, dt, dt, 50)
                    bk.run(dt, n_steps, 0, 0, flags=L.RUN_WRITE_STATE, want_dyn=False, work=bk.work,
                           t_main=bk.stamps[timed_step[0]] if record else None)
            for bk in runs:
                cur.wait_stream(bk.stream)
        else:
            hb = runs[0]
            hb.bk.q.copy_(hb.q0); hb.bk.v.copy_(hb.v0); hb.eps_pi.copy_(hb.ep0); hb.hs.copy_(hb.hs0)
            hb.run(dt, n_steps, work=hb.work, t_main=hb.stamps[timed_step[0]] if record else None)
        if record:
            timed_step[0] += 1

    for _ in range(args.warmup if n_steps <= 1000 else 1):
        step()
    _barrier(torch, dist, world)
    mark0 = sampler.mark() if sampler is not None else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step(record=True)
    e1.record()
    _barrier(torch, dist, world)
    mark1 = sampler.mark() if sampler is not None else 0
    t_dev = e0.elapsed_time(e1) * 1e-3
    # e2e: pinned host state in, final host state out, every step
    host = []
    for r in runs:
        bk = r if c4 else r.bk
        host.append((r.q0.cpu().pin_memory(), r.v0.cpu().pin_memory(),
                     torch.empty(bk.q.shape, dtype=torch.float64).pin_memory(),
                     torch.empty(bk.v.shape, dtype=torch.float64).pin_memory()))
    h2d = sum(a.numel() * 8 + b.numel() * 8 for a, b, _, _ in host)
    d2h = h2d

    def step_e2e():
        for r, (q0h, v0h, qh, vh) in zip(runs, host):
            r.q0.copy_(q0h, non_blocking=True); r.v0.copy_(v0h, non_blocking=True)
        step()
        for r, (q0h, v0h, qh, vh) in zip(runs, host):
            bk = r if c4 else r.bk
            qh.copy_(bk.q, non_blocking=True); vh.copy_(bk.v, non_blocking=True)
        torch.cuda.synchronize()

    e2e_steps = steps if n_steps <= 1000 else 1
    if n_steps <= 1000:
        step_e2e()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    t_e2e = (time.perf_counter() - t0) * steps / e2e_steps
    t_dev, t_e2e = _allmax(torch, dist, world, dev, t_dev, t_e2e)
    stat_words = torch.cat([(r if c4 else r.bk).status for r in runs])
    n_bad = int((stat_words != 0).sum())
    n_nonfinite = int(((stat_words & 1) != 0).sum())
    if rank != 0:
        return None
    # ---- roofline from counted work
    peak = L.peak_flops(0, local)
    tss = [r.stamps[:steps].cpu().numpy() for r in runs]
    win = float(np.mean([max(t[k, 1] for t in tss) - min(t[k, 0] for t in tss) for k in range(steps)])) * 1e-9
    if c4:
        iters = sum(float(bk.work[:, 0].sum()) for bk in runs)
        solves = sum(float(bk.work[:, 1].sum()) for bk in runs)
        # SURVEY.md 8d "whfast": 14 N(N-1) per full-force kick + 2 (N-1) K, K = 60 flops x Newton iterations (measured)
        fl = sum(float(bk.B) * n_steps * 14.0 * bk.N * (bk.N - 1) for bk in runs) + 60.0 * iters
        counted = {"kepler_solves": solves, "mean_newton_iterations_executed": iters / max(solves, 1.0),
                   "flop_model": "SURVEY.md 8d: 14 N(N-1) + 60 x Newton iterations per Kepler solve, 2(N-1) solves per step"}
        kernel = "ensemble_main_kernel<N=3..5, whfast> (3 concurrent launches per step)"
    else:
        hb = runs[0]
        sweeps = float(hb.work[:, 0].sum())
        halves = float(hb.work[:, 1].sum())
        N = hb.bk.N
        # SURVEY.md 8d "ham_soft sub-step": 2 x 17 N(N-1) + (8N+4) S, S = sweeps x N(N-1) x (12 flops + 1 exp)
        fl = 0.5 * halves * 2 * 17.0 * N * (N - 1) + sweeps * N * (N - 1) * 13.0
        counted = {"s_half_flows": halves, "mean_jacobi_sweeps_per_eps_star_evaluation": sweeps / max(halves * (4 * N + 1), 1.0),
                   "flop_model": "SURVEY.md 8d: 2 x 17 N(N-1) per sub-step + 13 flops per pair and Jacobi sweep, "
                                 "4N+1 eps* evaluations per S half-flow (exp counted as 1 flop)"}
        kernel = "hamsoft_run_kernel<3>"
    ach = fl / win * 1e-12
    roof = {"bound": "fp64", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "frac_nominal": ach / NOMINAL_FP64, "peak_nominal": NOMINAL_FP64, "flops_per_step": fl, "ms": win * 1e3,
            "traffic": _traffic("c4_main_bytes_per_step" if c4 else "c1_hamsoft_bytes_per_launch"),
            "timing": "%globaltimer stamps published by the run kernels inside the timed steps", "counted_work": counted,
            "note": "algorithmic flops per SURVEY.md 8d with divisions / sqrt / exp counted as ONE flop each although each "
                    "costs 10-30 FP64 instructions: the FP64-pipe utilisation of these kernels is in profiles/ (ncu)",
            "share_of_step": win / (t_dev / steps)}
    sys_steps = float(B) * world * n_steps * steps
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        pool = CpuPool(cores)
        if c4:
            inp = _c4_inputs(3 * cores * 8, 7)
            jobs = [(inp[N][0][i], inp[N][1][i], inp[N][2][i], 2000, dt) for N in sorted(inp) for i in range(inp[N][0].shape[0])]
            rate, dtc = pool.run(_cpu_c4, jobs)
            sample = f"{len(jobs)} planetary systems x 2000 whfast steps in {cores} processes, {dtc:.1f} s"
        else:
            m, q, v = _c1_inputs(6 * cores, 7)
            jobs = [(m[i], q[i], v[i], 1000, dt) for i in range(6 * cores)]
            rate, dtc = pool.run(_cpu_c1, jobs)
            sample = f"{6 * cores} jittered README systems x 1000 ham_soft steps in {cores} processes, {dtc:.1f} s"
        pool.close()
        cpu = {"value": rate, "unit": "system-steps/s", "cores": cores, "kind": cpu_kind(), "sample": sample}
    name = ("C4 WHFast + Kepler planetary ensemble (BASELINE.json configs[3]): star + 2-4 planets near 3:2/2:1/5:3, "
            "half TTV cohort, dt = 0.01 x 2 pi, bug-compatible Kepler solver") if c4 else \
           ("C1 README 3-body ham_soft (BASELINE.json configs[0]) batched: jittered copies, dt = 0.01, adaptive-epsilon "
            "Strang flow with the finite-difference eps* gradient")
    line = {
        "metric": "system-steps/s", "value": sys_steps / t_dev, "unit": "system-steps/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "systems_per_gpu": B, "integrator_steps_per_bench_step": n_steps,
                   "buckets": {str((r if c4 else r.bk).N): int((r if c4 else r.bk).B) for r in runs}},
        "e2e": {"value": sys_steps / t_e2e, "unit": "system-steps/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t_e2e / steps},
        "gpu_launches": int((2 * len(runs) if c4 else 1) * steps), "roofline": roof, "cpu_baseline": cpu,
        "checks": {"systems_with_nonzero_status": n_bad, "systems_nonfinite": n_nonfinite},
    }
    if not c4:
        line["api_single_system"] = _c1_api_single_system(1000)
    if standalone:
        line["clocks"] = sampler.stop(mark0, mark1) if sampler is not None else None
    return line


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
        return None

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
        pool = CpuPool(cores)
        rate, dtc = pool.run(_cpu_c2, jobs)
        pool.close()
        cpu = {"value": rate, "unit": "system-steps/s", "cores": cores, "kind": "port",
               "sample": f"the same 10 systems, oracle run_stability_analysis in {cores} processes, {dtc:.1f} s"}
    return ({
        "metric": "system-steps/s", "value": sys_steps / t, "unit": "system-steps/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 MLTrainingPipeline.quick_test cohort (BASELINE.json configs[1]): 10 systems N=3,4,5, verlet, "
                               "1000 steps + MEGNO, mode full, through NBodySimulation / BatchStabilityAnalyzer (includes "
                               "building the 10 simulation objects and the DataFrame)",
                   "columns": int(df.shape[1])},
        "e2e": {"value": sys_steps / t, "unit": "system-steps/s", "h2d_bytes_per_step": int(10 * 4 * 5 * 8 * 3),
                "d2h_bytes_per_step": int(10 * 47 * 8), "api": "BatchStabilityAnalyzer.analyze_batch"},
        "gpu_launches": None, "roofline": None, "cpu_baseline": cpu})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 20; 5 for c4 / c1 / largen)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ensemble", choices=["ensemble", "largen", "c4", "c1", "c2"])
    ap.add_argument("--systems", type=int, default=1 << 20, help="systems per GPU (weak scaling)")
    ap.add_argument("--particles", "--n", dest="n", type=int, default=1 << 20, help="particles of the large-N system")
    ap.add_argument("--n-hamsoft", type=int, default=0, dest="n_hamsoft",
                    help="particles for the ham_soft Strang sub-step timing of --workload largen (default: --n)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-largen", action="store_true", help="skip the large-N section of the default line")
    ap.add_argument("--no-largen-hamsoft", action="store_true", dest="no_largen_hamsoft",
                    help="skip the large-N ham_soft Strang sub-step timing of the default line")
    ap.add_argument("--no-secondary", action="store_true", dest="no_secondary",
                    help="skip the C1 / C4 sections of the default line")
    ap.add_argument("--heavy-threshold", type=int, default=-1, dest="heavy_threshold",
                    help="tuning sweeps only: fixed n_sub threshold of the latency mappings on the device-resident path "
                         "(-1 = the automatic N-only rule, the product setting)")
    ap.add_argument("--sets", type=int, default=1,
                    help="device path: buffer / stream sets that consecutive steps alternate between (steps in flight)")
    ap.add_argument("--order", default="", help="tuning only: launch order of the N buckets on the device path, e.g. 8,7,6,5,4,3")
    ap.add_argument("--horizon", type=int, default=0,
                    help="integrator steps per system for --workload c4 / c1 (default 1000; C4 as worded: 1000000)")
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
