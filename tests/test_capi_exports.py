"""CPU: the C-ABI library loads and exports every symbol include/nbody_b200.h declares (no compute)."""
import os
import re

from conftest import ROOT


def test_exports_match_header():
    from nbodysimproject_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "nbody_b200.h")).read()
    declared = set(re.findall(r"\b(nb_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.nb_version() >= 100


def test_header_constants_match_python():
    from nbodysimproject_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "nbody_b200.h")).read()
    def enum_block(first):
        i = hdr.index(first)
        blk = hdr[hdr.rindex("enum {", 0, i):hdr.index("};", i)]
        blk = re.sub(r"/\*.*?\*/", "", blk, flags=re.S)
        return [t.strip().split("=")[0].strip() for t in blk.replace("enum {", "").split(",") if t.strip()]
    dyn = enum_block("NB_F_IS_STABLE")
    assert dyn[-1] == "NB_N_DYN" and len(dyn) - 1 == _lib.N_DYN
    stat = enum_block("NB_S_TOTAL_MASS")
    assert stat[-1] == "NB_N_STATIC" and len(stat) - 1 == _lib.N_STATIC
    hs = enum_block("NB_HS_K_SOFT")
    assert hs[-1] == "NB_HS_NPARAM" and len(hs) - 1 == _lib.N_HS


def test_no_cpu_fallback_without_gpu():
    import pytest
    import torch
    from nbodysimproject_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.NBodyB200Error):
        from nbodysimproject_b200 import ensemble
        ensemble.pair_batched([[[0.0, 0.0], [1.0, 0.0]]], [[1.0, 1.0]], 0.0)


def test_product_path_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package (nor the C sources) may import or call it."""
    import glob
    pkg = os.path.join(ROOT, "nbodysimproject_b200")
    files = glob.glob(os.path.join(pkg, "*.py")) + glob.glob(os.path.join(pkg, "csrc", "*.cu*"))
    assert len(files) > 15
    for f in files:
        src = open(f).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
        assert "oracle." not in src and "import oracle" not in src, f


def test_bad_arguments_return_error_codes_without_touching_the_gpu():
    """Every entry point validates its arguments first and reports NB_ERR_ARG / NB_ERR_UNSUPPORTED through the return
    code + nb_last_error(), never by throwing or crashing (include/nbody_b200.h conventions).  No CUDA call is reached."""
    import ctypes as C
    from nbodysimproject_b200 import _lib
    lib = _lib.load()
    N = None
    one = (C.c_double * 16)()
    p = C.cast(one, C.c_void_p)
    cases = [
        lib.nb_pair_batched_f64(N, N, N, 1.0, 4, 3, N, N, N, N),
        lib.nb_pair_batched_f64(p, p, p, 1.0, 4, 99, p, p, p, N),
        lib.nb_variational_batched_f64(N, N, N, N, 1.0, 4, 3, N, N),
        lib.nb_ensemble_run_f64(N, N, N, N, 1.0, 4, 3, 0, 0, 0.01, 1, 0, 0, N, N, N, N, N, N, N, N, N, N),
        lib.nb_ensemble_run_f64(p, p, p, p, 1.0, 4, 3, 0, 0, 0.01, 1, 0, 5, N, N, N, N, N, N, N, N, N, N),   # megno without draws
        lib.nb_ensemble_run_adaptive_f64(N, N, N, N, N, 1.0, 4, 3, 0, 0.01, 1, N, 1e9, 5, N, N, N, N),
        lib.nb_sort_by_nsub(N, 4, 3, N, N, -1, N),
        lib.nb_sort_by_nsub(p, 4, 3, p, p, 64, N),                                     # threshold out of range
        lib.nb_largeN_accel_f32(N, 10, 0, 10, 0.1, 1.0, N, N, N, -1, N),
        lib.nb_largeN_accel_f32(p, 10, 5, 10, 0.1, 1.0, p, N, p, -1, N),                # i-range outside n_total
        lib.nb_largeN_accel_f32(p, 10, 0, 10, 0.1, 1.0, p, N, N, -1, N),                # no workspace
        lib.nb_largeN_accel_f32(p, 10, 0, 10, 0.1, 1.0, p, N, p, 99, N),                # unknown variant
        lib.nb_largeN_pass_f32(7, p, N, 10, 0, 10, N, 0.0, p, N, N),                      # unknown kind
        lib.nb_largeN_pass_f32(0, p, N, 10, 0, 10, N, 0.0, p, N, N),                      # DENSITY without h
        lib.nb_mlp_classify_f32(p, p, p, 65, p, p, p, p, p, p, p, 0.0, 0.5, 4, p, N, N),   # F > 64
        lib.nb_generate_ensemble_f64(9, 3, 4, 1, 0, p, p, p, p, N),                     # unknown cohort
        lib.nb_generate_ensemble_f64(1, 4, 4, 1, 0, p, p, p, p, N),                     # hierarchical needs N = 3
        lib.nb_ensemble_analyze_host_async(p, p, p, p, 1.0, 4, 3, 0, 0, 0.01, 0.01, 0.01, 10, 0, 50, N, N, p, N, N, N, 0, 99),
        lib.nb_host_sync(-1),
    ]
    assert all(rc < 0 for rc in cases), cases
    assert lib.nb_last_error().decode() != ""
    # adaptive softening exists for verlet / yoshida4 only
    assert lib.nb_ensemble_run_adaptive_f64(p, p, p, p, p, 1.0, 4, 3, 2, 0.01, 1, N, 1e9, 5, N, N, N, N) == -3
