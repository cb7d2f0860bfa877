#!/bin/bash
python -m pytest tests/test_gpu_hamsoft.py tests/test_gpu_host_path.py tests/test_gpu_api.py tests/test_gpu_hamsoft_mid.py -x -q 2>&1 | tail -3
python tools/profile_hamsoft.py 65536 100 2>&1 | tail -5
