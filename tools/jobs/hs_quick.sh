#!/bin/bash
python -m pytest tests/test_gpu_hamsoft.py tests/test_gpu_api.py -x -q 2>&1 | tail -5
python tools/profile_hamsoft.py 131072 100 2>&1 | tail -8
