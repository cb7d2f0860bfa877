#!/bin/bash
( time python -m pytest tests/test_gpu_c4_horizon.py -x -q 2>&1 | tail -15 ) 2>&1
cat gpurun_out/c4_horizon_stats.json
