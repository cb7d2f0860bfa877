"""ctypes binding of libnbody_b200.so (the C ABI in include/nbody_b200.h).

There is no CPU fallback: if the shared library is missing, or a compute entry point is called
without a CUDA device, this module raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C nbodysimproject_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnbody_b200.so")

# ---- constants mirrored from include/nbody_b200.h ------------------------------------------------
NB_OK = 0
MODE_VERLET, MODE_YOSHIDA4, MODE_WHFAST, MODE_HAMSOFT = 0, 1, 2, 3
MODES = {"verlet": MODE_VERLET, "yoshida4": MODE_YOSHIDA4, "whfast": MODE_WHFAST, "ham_soft": MODE_HAMSOFT}
STATUS_OK, STATUS_NONFINITE, STATUS_EPS_OOB, STATUS_KEPLER_NOCONV = 0, 1, 2, 4
PREP_REMOVE_COM, PREP_CTOR_KICK, PREP_SNAPSHOT_KICK, PREP_STATIC_FEATURES = 1, 2, 4, 8
RUN_ENERGY, RUN_WRITE_STATE, RUN_KEPLER_EXACT = 1, 2, 4

DYN_COLUMNS = [
    "is_stable", "energy_drift", "angular_momentum_drift", "com_drift_mean", "com_drift_max",
    "j_eps_mean", "j_eps_std", "theta_eps_mean", "theta_eps_std", "cos_theta_mean", "cos_theta_min",
    "ang_mom_var_mean", "ang_mom_var_max", "tidal_trace_mean", "tidal_trace_max", "MEGNO", "lyapunov_time",
    "_E0", "_E1", "_L0", "_L1", "_t_end",
]
N_DYN = len(DYN_COLUMNS)
STATIC_COLUMNS = [
    "total_mass", "mass_variance", "mass_ratio_max", "mass_center_offset",
    "mean_separation", "std_separation", "min_separation", "max_separation", "separation_ratio",
    "mean_speed", "std_speed", "max_speed", "mean_relative_velocity", "max_relative_velocity",
    "kinetic_energy", "potential_energy", "total_energy", "virial_ratio", "energy_per_mass", "is_bound",
    "total_angular_momentum", "mean_specific_angular_momentum", "angular_momentum_variance",
    "softening_mean", "softening_std",
]
N_STATIC = len(STATIC_COLUMNS)
HS_PARAMS = ["k_soft", "mu_soft", "eps_min", "eps_max", "alpha_run", "k_wall", "barrier_n", "eta",
             "j_max_cap", "lambda", "policy", "theta_imp", "theta_cap", "chi_pi", "omega_spr0", "s0", "flags"]
HS_FLAG_FREEZE_S, HS_FLAG_S_ONLY = 1, 2
N_HS = len(HS_PARAMS)

EXPORTS = [
    "nb_last_error", "nb_version", "nb_pair_batched_f64", "nb_variational_batched_f64",
    "nb_ensemble_prepare_f64", "nb_ensemble_run_f64", "nb_ensemble_run_counted_f64", "nb_sort_by_nsub",
    "nb_ensemble_run_adaptive_f64", "nb_ensemble_analyze_adaptive_f64", "nb_ensemble_analyze_host",
    "nb_ensemble_analyze_host_async", "nb_ensemble_analyze_host_ex", "nb_host_sync", "nb_generate_tangent_f64", "nb_hamsoft_setup_f64", "nb_hamsoft_probe_f64",
    "nb_largeN_accel_f32", "nb_largeN_kick_drift_f32", "nb_largeN_pass_f32", "nb_largeN_tile_boxes_f32", "nb_mlp_classify_f32", "nb_generate_ensemble_f64",
    "nb_peak_flops",
]

HOST_COMPACT_DYN, HOST_KEEP_V, HOST_DEVICE_TANGENT, HOST_ADAPTIVE, HOST_HS_NO_CALIBRATE = 1, 2, 4, 8, 16
N_DYN_USER = 17


class HostOpts(C.Structure):
    """nb_host_opts (include/nbody_b200.h)."""
    _fields_ = [("size", C.c_uint32), ("flags", C.c_uint32), ("n_chunks", C.c_int32), ("barrier_exponent", C.c_int32),
                ("tangent_seed", C.c_uint64), ("first_index", C.c_uint64), ("hs_params", C.c_void_p),
                ("eps_pi", C.c_void_p), ("soft_par", C.c_void_p), ("eps_start", C.c_void_p), ("eps_energy", C.c_void_p),
                ("energy_delta", C.c_void_p), ("k_wall", C.c_double)]

    def __init__(self, **kw):
        super().__init__()
        self.size = C.sizeof(HostOpts)
        self._keep = []
        for k, val in kw.items():
            if hasattr(val, "ctypes") or hasattr(val, "data_ptr"):
                self._keep.append(val)                      # keep the buffer alive as long as the struct
                val = val.data_ptr() if hasattr(val, "data_ptr") else val.ctypes.data
            setattr(self, k, val)


_lib = None


class NBodyB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (no GPU needed to load it; compute calls need one)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NBodyB200Error(
            f"{LIB_PATH} is missing: the CUDA library has not been built. There is no CPU fallback; "
            "run __graft_entry__.build() or `make -C nbodysimproject_b200/csrc`.")
    lib = C.CDLL(LIB_PATH)
    p, i, d, u, f = C.c_void_p, C.c_int, C.c_double, C.c_uint, C.c_float
    lib.nb_last_error.restype = C.c_char_p
    lib.nb_last_error.argtypes = []
    lib.nb_version.restype = i
    lib.nb_pair_batched_f64.argtypes = [p, p, p, d, i, i, p, p, p, p]
    lib.nb_variational_batched_f64.argtypes = [p, p, p, p, d, i, i, p, p]
    lib.nb_ensemble_prepare_f64.argtypes = [p, p, p, p, d, i, i, i, u, d, d, d, i, p, p, p, p]
    lib.nb_ensemble_run_f64.argtypes = [p, p, p, p, d, i, i, i, u, d, i, i, i, p, p, p, p, p, p, p, p, p, p]
    lib.nb_ensemble_run_counted_f64.argtypes = [p, p, p, p, d, i, i, i, u, d, i, i, i, p, p, p, p, p, p, p, p, p, p, p, p]
    lib.nb_sort_by_nsub.argtypes = [p, i, i, p, p, i, p]
    lib.nb_ensemble_run_adaptive_f64.argtypes = [p, p, p, p, p, d, i, i, i, d, i, p, d, i, p, p, p, p]
    lib.nb_ensemble_analyze_adaptive_f64.argtypes = [p, p, p, p, p, p, d, i, i, i, d, i, i, i, p, p, p, d, i, p, p, p, p]
    lib.nb_ensemble_analyze_host.argtypes = [p, p, p, p, d, i, i, i, u, d, d, d, i, i, i, p, p, p, p, p, p, i]
    lib.nb_ensemble_analyze_host_async.argtypes = lib.nb_ensemble_analyze_host.argtypes + [i]
    lib.nb_ensemble_analyze_host_ex.argtypes = lib.nb_ensemble_analyze_host.argtypes + [i, p]
    lib.nb_generate_tangent_f64.argtypes = [i, i, C.c_uint64, C.c_uint64, p, p, p]
    lib.nb_host_sync.argtypes = [i]
    lib.nb_hamsoft_setup_f64.argtypes = [p, p, d, i, i, u, d, p, p, p, p]
    lib.nb_hamsoft_probe_f64.argtypes = [p, p, p, d, i, i, p, p, p, p]
    lib.nb_largeN_accel_f32.argtypes = [p, i, i, i, f, f, p, p, p, i, p]
    lib.nb_largeN_kick_drift_f32.argtypes = [p, p, p, i, f, f, p]
    lib.nb_largeN_pass_f32.argtypes = [i, p, p, i, i, i, p, f, p, p, p]
    lib.nb_largeN_tile_boxes_f32.argtypes = [p, p, i, p, p]
    lib.nb_mlp_classify_f32.argtypes = [p, p, p, i, p, p, p, p, p, p, p, f, f, i, p, p, p]
    lib.nb_generate_ensemble_f64.argtypes = [i, i, i, C.c_uint64, C.c_uint64, p, p, p, p, p]
    lib.nb_peak_flops.argtypes = [i, i, C.POINTER(C.c_double)]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("nb_last_error",):
            fn.restype = i
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != NB_OK:
        msg = load().nb_last_error().decode("utf-8", "replace")
        raise NBodyB200Error(f"{what or 'nbody_b200'} failed with code {rc}: {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise NBodyB200Error("no CUDA device: nbodysimproject_b200 has no CPU fallback (B200 / sm_100a only)")
    return torch


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def peak_flops(which: int, device: int = 0) -> float:
    require_cuda()
    out = C.c_double(0.0)
    check(load().nb_peak_flops(which, device, C.byref(out)), "nb_peak_flops")
    return out.value
