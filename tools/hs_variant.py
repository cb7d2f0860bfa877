"""Tuning helper: tools/profile_hamsoft.py with another build of the shared library.  usage: python tools/hs_variant.py lib.so [B] [steps]"""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodysimproject_b200 import _lib as L  # noqa: E402

L.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = ["profile_hamsoft.py"] + (sys.argv[2:] or ["65536", "100"])
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profile_hamsoft.py"), run_name="__main__")
