#!/bin/bash
# mid-N ham_soft parity + the existing ham_soft / mid-N / host-path tests
python -m pytest tests/test_gpu_hamsoft_mid.py tests/test_gpu_midn.py tests/test_gpu_hamsoft.py tests/test_gpu_host_path.py -x -q 2>&1 | tail -25
