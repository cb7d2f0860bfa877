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
