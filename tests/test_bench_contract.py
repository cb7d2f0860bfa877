"""CPU: the reference arm of bench.py (the only arm that runs without a GPU) prints one JSON line with the contract's
keys, and the B200 arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


@pytest.mark.timeout(400)
def test_reference_arm_lines():
    for extra in ([], ["--workload", "largen"]):
        r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", *extra)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        assert len(lines) == 1
        d = json.loads(lines[0])
        for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                    "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
            assert key in d, key
        assert d["impl"] == "reference" and d["value"] > 0 and d["gpu_launches"] == 0
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
        assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
        assert "workload" in d["config"]


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run("--steps", "1", "--warmup", "3")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
