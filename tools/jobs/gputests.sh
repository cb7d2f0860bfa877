#!/bin/bash
# full GPU suite with per-test lines written as they finish (a timeout keeps the partial log)
mkdir -p gpurun_out
( time stdbuf -oL python -m pytest tests -m gpu -x -v --durations=15 ) 2>&1 | stdbuf -oL grep -v "^$" > gpurun_out/pytest_gpu_r2.log
tail -30 gpurun_out/pytest_gpu_r2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
