"""Batched ensemble driver: NumPy / torch buffers in, C-ABI kernels, feature tables out.

This is the host-side mirror of the reference's per-system Python loops
(batch_stability_analyzer.py:62-80 -> stability_analyzer.py:69-259 -> simulation.py:667-676): systems are
bucketed by (N, integrator mode, G), each bucket becomes one prepare + sort + run launch sequence.
PyTorch is only used for device memory and streams.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib as L


def _dev(device=None):
    torch = L.require_cuda()
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _to_dev(a, dtype, device):
    torch = L.require_cuda()
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype).contiguous()
    a = np.ascontiguousarray(a)
    if not a.flags.writeable:            # e.g. np.broadcast_to views, npz members: torch wants a writable buffer
        a = a.copy()
    return torch.as_tensor(a, dtype=dtype).to(device)


# ---------------------------------------------------------------------------------------------
# a1-a6 stand-alone batched pair kernels
# ---------------------------------------------------------------------------------------------

def pair_batched(q, m, eps, G: float = 1.0, device=None, want_acc=True, want_U=True, want_dV=True):
    """acc[B,N,2], U[B], dV/deps[B] for B systems of N bodies (forces.py:63-112, potential.py:23-64)."""
    torch = L.require_cuda()
    device = _dev(device)
    q = _to_dev(q, torch.float64, device)
    m = _to_dev(m, torch.float64, device)
    B, N = int(m.shape[0]), int(m.shape[1])
    eps = _to_dev(np.broadcast_to(np.asarray(eps, dtype=np.float64), (B,)) if not isinstance(eps, torch.Tensor) else eps,
                  torch.float64, device)
    acc = torch.empty((B, N, 2), dtype=torch.float64, device=device) if want_acc else None
    U = torch.empty((B,), dtype=torch.float64, device=device) if want_U else None
    dV = torch.empty((B,), dtype=torch.float64, device=device) if want_dV else None
    with torch.cuda.device(device):
        L.check(L.load().nb_pair_batched_f64(L.ptr(q), L.ptr(m), L.ptr(eps), float(G), B, N, L.ptr(acc), L.ptr(U),
                                             L.ptr(dV), L.stream_ptr()), "nb_pair_batched_f64")
    return acc, U, dV


def variational_batched(q, m, s2, dr, G: float = 1.0, device=None):
    """TangentMap.variational_accel for B systems (tangent_map.py:21-59)."""
    torch = L.require_cuda()
    device = _dev(device)
    q = _to_dev(q, torch.float64, device)
    m = _to_dev(m, torch.float64, device)
    dr = _to_dev(dr, torch.float64, device)
    B, N = int(m.shape[0]), int(m.shape[1])
    s2 = _to_dev(np.broadcast_to(np.asarray(s2, dtype=np.float64), (B,)) if not isinstance(s2, torch.Tensor) else s2,
                 torch.float64, device)
    da = torch.empty((B, N, 2), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        L.check(L.load().nb_variational_batched_f64(L.ptr(q), L.ptr(m), L.ptr(s2), L.ptr(dr), float(G), B, N,
                                                    L.ptr(da), L.stream_ptr()), "nb_variational_batched_f64")
    return da


# ---------------------------------------------------------------------------------------------
# one (N, mode, G) bucket resident on the device
# ---------------------------------------------------------------------------------------------

@dataclass
class BucketResult:
    dyn: np.ndarray              # [B, N_DYN]
    static: Optional[np.ndarray]  # [B, N_STATIC] or None
    n_sub: np.ndarray            # [B]
    status: np.ndarray           # [B]
    v_kicked: Optional[np.ndarray]
    q_final: Optional[np.ndarray] = None
    v_final: Optional[np.ndarray] = None
    extra: Optional[dict] = None


class DeviceBucket:
    """Device-resident state of B systems with the same N / mode / G (torch tensors as buffers)."""

    def __init__(self, m, q, v, eps, G=1.0, mode="verlet", device=None):
        torch = L.require_cuda()
        self.torch = torch
        self.device = _dev(device)
        self.mode = L.MODES[mode] if isinstance(mode, str) else int(mode)
        self.G = float(G)
        self.m = _to_dev(m, torch.float64, self.device)
        self.q = _to_dev(q, torch.float64, self.device).clone()
        self.v = _to_dev(v, torch.float64, self.device).clone()
        self.B, self.N = int(self.m.shape[0]), int(self.m.shape[1])
        if not (2 <= self.N <= 64):
            raise L.NBodyB200Error(f"ensemble kernels support N = 2..64 bodies per system, got {self.N}")
        self.eps = _to_dev(np.broadcast_to(np.asarray(eps, dtype=np.float64), (self.B,))
                           if not isinstance(eps, torch.Tensor) else eps, torch.float64, self.device)
        self.n_sub = torch.ones((self.B,), dtype=torch.int32, device=self.device)
        self.h_sub_ref = torch.empty((self.B,), dtype=torch.float64, device=self.device)
        self.perm = None
        self.static = None
        self.status = torch.zeros((self.B,), dtype=torch.int32, device=self.device)
        self._bins = torch.zeros((128,), dtype=torch.int32, device=self.device)

    def prepare(self, flags: int, kick_dt: float, sched_dt: float, dt: float, split_n_max: int = 50,
                want_static: bool = False):
        torch = self.torch
        if want_static:
            self.static = torch.empty((self.B, L.N_STATIC), dtype=torch.float64, device=self.device)
            flags |= L.PREP_STATIC_FEATURES
        with torch.cuda.device(self.device):
            L.check(L.load().nb_ensemble_prepare_f64(
                L.ptr(self.m), L.ptr(self.q), L.ptr(self.v), L.ptr(self.eps), self.G, self.B, self.N, self.mode,
                int(flags), float(kick_dt), float(sched_dt), float(dt), int(split_n_max), L.ptr(self.h_sub_ref),
                L.ptr(self.n_sub), L.ptr(self.static), L.stream_ptr()), "nb_ensemble_prepare_f64")

    def set_n_sub_from_h(self, h_sub_ref, dt, split_n_max=50):
        """integrator.py:86-92 with an externally frozen h_sub_ref (NBodySimulation.step)."""
        torch = self.torch
        h = _to_dev(h_sub_ref, torch.float64, self.device)
        n = torch.clamp(torch.ceil(abs(float(dt)) / h), 1, int(split_n_max)).to(torch.int32)
        self.n_sub = n.contiguous()

    def sort(self, heavy_threshold: int = -1):
        """n_sub-descending permutation; heavy_threshold -1 = the automatic, N-only rule (product setting)."""
        torch = self.torch
        self.perm = torch.empty((self.B,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.load().nb_sort_by_nsub(L.ptr(self.n_sub), self.B, self.N, L.ptr(self.perm), L.ptr(self._bins),
                                             int(heavy_threshold), L.stream_ptr()), "nb_sort_by_nsub")

    def run(self, dt, n_steps, sample_interval=0, n_megno=0, raw_dr=None, raw_dv=None, flags=0, want_dyn=True,
            eps_pi=None, hs_params=None, work=None, t_main=None):
        """`t_main` = int64 device tensor [2] initialised to {2^63-1, 0}: earliest start / latest end of the main-phase
        kernels in %globaltimer ns.  `work` (optional float64 [B, 2] device tensor) receives the counted work of the run: whfast {Newton
        iterations, Kepler solves}, ham_soft {Jacobi sweeps, S half-flows} (nb_ensemble_run_counted_f64)."""
        torch = self.torch
        dyn = torch.empty((self.B, L.N_DYN), dtype=torch.float64, device=self.device) if want_dyn else None
        rdr = _to_dev(raw_dr, torch.float64, self.device) if n_megno > 0 else None
        rdv = _to_dev(raw_dv, torch.float64, self.device) if n_megno > 0 else None
        with torch.cuda.device(self.device):
            L.check(L.load().nb_ensemble_run_counted_f64(
                L.ptr(self.m), L.ptr(self.q), L.ptr(self.v), L.ptr(self.eps), self.G, self.B, self.N, self.mode,
                int(flags), float(dt), int(n_steps), int(sample_interval), int(n_megno), L.ptr(self.n_sub),
                L.ptr(self.perm), L.ptr(self._bins[64:]) if self.perm is not None else None, L.ptr(rdr), L.ptr(rdv),
                L.ptr(eps_pi), L.ptr(hs_params), L.ptr(dyn), L.ptr(self.status), L.ptr(work), L.ptr(t_main), L.stream_ptr()),
                "nb_ensemble_run_counted_f64")
        return dyn


def analysis_plan(n_steps: int, mode: str):
    """stability_analyzer.py:69-141: (sample_interval, n_megno) for an analysis mode."""
    n_steps = max(1, int(n_steps))
    if mode == "minimal":
        return 0, 0
    interval = max(1, n_steps // 100)
    n_megno = 0
    if mode == "full":
        n_samp = min(50, n_steps // 2)
        n_megno = min(100, n_samp) if n_samp > 0 else 0
    return interval, n_megno


def analyze_bucket(m, q, v, eps, G=1.0, mode="verlet", n_steps=1000, dt=0.01, analysis_mode="full",
                   raw_dr=None, raw_dv=None, prep_flags=L.PREP_SNAPSHOT_KICK, kick_dt=0.01, sched_dt=0.01,
                   split_n_max=50, device=None, via="device") -> BucketResult:
    """run_stability_analysis for B same-N systems.  `via="host"` goes through the single C-ABI call
    nb_ensemble_analyze_host (host buffers in, host buffers out); `via="device"` drives the device-pointer
    entry points with torch tensors."""
    m = np.ascontiguousarray(m, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    v = np.array(v, dtype=np.float64, order="C")
    B, N = m.shape
    eps = np.ascontiguousarray(np.broadcast_to(np.asarray(eps, dtype=np.float64), (B,)))
    n_steps = max(1, int(n_steps))
    interval, n_megno = analysis_plan(n_steps, analysis_mode)
    want_static = analysis_mode == "full"
    imode = L.MODES[mode]
    if via == "host":
        torch = L.require_cuda()
        dev_index = _dev(device).index or 0
        dyn = np.empty((B, L.N_DYN))
        stat = np.empty((B, L.N_STATIC)) if want_static else None
        nsub = np.empty((B,), dtype=np.int32)
        status = np.empty((B,), dtype=np.int32)
        rdr = np.ascontiguousarray(raw_dr, dtype=np.float64) if n_megno > 0 else None
        rdv = np.ascontiguousarray(raw_dv, dtype=np.float64) if n_megno > 0 else None
        if analysis_mode == "minimal":
            raise L.NBodyB200Error("via='host' implements the core/full analysis")
        L.check(L.load().nb_ensemble_analyze_host(
            L.ptr(m), L.ptr(q), L.ptr(v), L.ptr(eps), float(G), B, N, imode, int(prep_flags), float(kick_dt),
            float(sched_dt), float(dt), n_steps, n_megno, int(split_n_max), L.ptr(rdr), L.ptr(rdv), L.ptr(dyn),
            L.ptr(stat), L.ptr(nsub), L.ptr(status), dev_index), "nb_ensemble_analyze_host")
        return BucketResult(dyn, stat, nsub, status, v)
    bk = DeviceBucket(m, q, v, eps, G, mode, device)
    bk.prepare(prep_flags, kick_dt, sched_dt, dt, split_n_max, want_static)
    v_kicked = bk.v.cpu().numpy() if prep_flags & (L.PREP_REMOVE_COM | L.PREP_CTOR_KICK | L.PREP_SNAPSHOT_KICK) else None
    bk.sort()
    dyn = bk.run(dt, n_steps, interval, n_megno, raw_dr, raw_dv, flags=L.RUN_ENERGY)
    return BucketResult(dyn.cpu().numpy(), bk.static.cpu().numpy() if want_static else None,
                        bk.n_sub.cpu().numpy(), bk.status.cpu().numpy(), v_kicked)


def analyze_host(m, q, v, eps, G=1.0, mode="verlet", n_steps=1000, dt=0.01, n_megno=0, raw_dr=None, raw_dv=None,
                 prep_flags=L.PREP_SNAPSHOT_KICK, kick_dt=0.01, sched_dt=0.01, split_n_max=50, device=None, slot=0,
                 want_static=True, compact=False, keep_v=False, tangent_seed=None, first_index=0, n_chunks=0,
                 hs_params=None, eps_pi=None, soft_par=None, eps_start=None, eps_energy=None, k_wall=0.0,
                 barrier_exponent=0, calibrate=True) -> BucketResult:
    """One nb_ensemble_analyze_host_ex call (HOST buffers in and out; chunk-pipelined H2D -> kernels -> D2H).
    mode "ham_soft" runs the reference's default integrator (constructor calibration + analysis in the same call);
    `soft_par` switches on classic adaptive softening.  `v` is mutated like the reference's snapshot() unless keep_v.
    Returns BucketResult with `extra` = dict(eps_pi=..., energy_delta=...) where applicable."""
    torch = L.require_cuda()
    dev_index = _dev(device).index or 0
    m = np.ascontiguousarray(m, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    if not (isinstance(v, np.ndarray) and v.dtype == np.float64 and v.flags.c_contiguous and v.flags.writeable):
        v = np.array(v, dtype=np.float64, order="C")
    B, N = m.shape
    eps = np.ascontiguousarray(np.broadcast_to(np.asarray(eps, dtype=np.float64), (B,)))
    flags = (L.HOST_COMPACT_DYN if compact else 0) | (L.HOST_KEEP_V if keep_v else 0)
    kw = dict(n_chunks=int(n_chunks), first_index=int(first_index), k_wall=float(k_wall),
              barrier_exponent=int(barrier_exponent))
    if tangent_seed is not None:
        flags |= L.HOST_DEVICE_TANGENT
        kw["tangent_seed"] = int(tangent_seed)
    if not calibrate:
        flags |= L.HOST_HS_NO_CALIBRATE
    ep = ed = None
    if hs_params is not None:
        kw["hs_params"] = np.ascontiguousarray(hs_params, dtype=np.float64)
    if eps_pi is not None:
        ep = np.array(eps_pi, dtype=np.float64, order="C")
        kw["eps_pi"] = ep
    if soft_par is not None:
        flags |= L.HOST_ADAPTIVE
        kw["soft_par"] = np.ascontiguousarray(soft_par, dtype=np.float64)
        ed = np.zeros((B,))
        kw["energy_delta"] = ed
        if eps_start is not None:
            kw["eps_start"] = np.ascontiguousarray(np.broadcast_to(np.asarray(eps_start, dtype=np.float64), (B,)))
        if eps_energy is not None:
            kw["eps_energy"] = np.ascontiguousarray(np.broadcast_to(np.asarray(eps_energy, dtype=np.float64), (B,)))
    opts = L.HostOpts(flags=flags, **kw)
    dyn = np.empty((B, L.N_DYN_USER if compact else L.N_DYN))
    stat = np.empty((B, L.N_STATIC)) if want_static else None
    nsub = np.empty((B,), dtype=np.int32)
    status = np.empty((B,), dtype=np.int32)
    rdr = np.ascontiguousarray(raw_dr, dtype=np.float64) if (n_megno > 0 and tangent_seed is None) else None
    rdv = np.ascontiguousarray(raw_dv, dtype=np.float64) if (n_megno > 0 and tangent_seed is None) else None
    import ctypes
    lib = L.load()
    L.check(lib.nb_ensemble_analyze_host_ex(
        L.ptr(m), L.ptr(q), L.ptr(v), L.ptr(eps), float(G), B, N, L.MODES[mode] if isinstance(mode, str) else int(mode),
        int(prep_flags), float(kick_dt), float(sched_dt), float(dt), int(n_steps), int(n_megno), int(split_n_max),
        L.ptr(rdr), L.ptr(rdv), L.ptr(dyn), L.ptr(stat), L.ptr(nsub), L.ptr(status), dev_index, int(slot),
        ctypes.byref(opts)), "nb_ensemble_analyze_host_ex")
    L.check(lib.nb_host_sync(int(slot)), "nb_host_sync")
    r = BucketResult(dyn, stat, nsub, status, v)
    r.extra = dict(eps_pi=ep, energy_delta=ed)
    return r


def advance_bucket(m, q, v, eps, h_sub_ref, G=1.0, mode="verlet", dt=0.01, n_steps=1, split_n_max=50, device=None,
                   kepler_exact=False):
    """n_steps x NBodySimulation.step(dt) for B same-N systems with frozen h_sub_ref (integrator.py:78-104)."""
    bk = DeviceBucket(m, q, v, eps, G, mode, device)
    bk.set_n_sub_from_h(h_sub_ref, dt, split_n_max)
    flags = L.RUN_WRITE_STATE | (L.RUN_KEPLER_EXACT if kepler_exact else 0)
    bk.run(dt, n_steps, 0, 0, flags=flags, want_dyn=False)
    return bk.q.cpu().numpy(), bk.v.cpu().numpy(), bk.status.cpu().numpy()


def advance_bucket_adaptive(m, q, v, eps, s0, min_softening, softening_scale, h_sub_ref, G=1.0, mode="verlet", dt=0.01,
                            n_steps=1, split_n_max=50, k_wall=1.0e9, barrier_exponent=5, energy_delta=None, device=None):
    """n_steps x NBodySimulation(adaptive_softening=True).step(dt) for B same-N systems (classic adaptive softening,
    softening_manager.py:298-336, 423-471, 541-547): returns q, v, eps_after_each_step[B, n_steps],
    softening_energy_delta[B], status[B]."""
    torch = L.require_cuda()
    dev = _dev(device)
    m_d = _to_dev(m, torch.float64, dev)
    q_d = _to_dev(q, torch.float64, dev).clone()
    v_d = _to_dev(v, torch.float64, dev).clone()
    B, N = int(m_d.shape[0]), int(m_d.shape[1])
    bc = lambda x: np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (B,)))
    eps_d = _to_dev(bc(eps), torch.float64, dev).clone()
    par = _to_dev(np.stack([bc(s0), bc(min_softening), bc(softening_scale)], 1), torch.float64, dev)
    h = _to_dev(bc(h_sub_ref), torch.float64, dev)
    n_sub = torch.clamp(torch.ceil(abs(float(dt)) / h), 1, int(split_n_max)).to(torch.int32).contiguous()
    e_d = _to_dev(bc(0.0 if energy_delta is None else energy_delta), torch.float64, dev).clone()
    hist = torch.empty((B, int(n_steps)), dtype=torch.float64, device=dev)
    status = torch.zeros((B,), dtype=torch.int32, device=dev)
    imode = L.MODES[mode] if isinstance(mode, str) else int(mode)
    with torch.cuda.device(dev):
        L.check(L.load().nb_ensemble_run_adaptive_f64(
            L.ptr(m_d), L.ptr(q_d), L.ptr(v_d), L.ptr(eps_d), L.ptr(par), float(G), B, N, imode, float(dt), int(n_steps),
            L.ptr(n_sub), float(k_wall), int(barrier_exponent), L.ptr(e_d), L.ptr(hist), L.ptr(status), L.stream_ptr()),
            "nb_ensemble_run_adaptive_f64")
    return (q_d.cpu().numpy(), v_d.cpu().numpy(), hist.cpu().numpy(), e_d.cpu().numpy(), status.cpu().numpy())
