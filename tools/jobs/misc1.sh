#!/bin/bash
python -m pytest tests/test_gpu_largen_hamsoft.py tests/test_gpu_hamsoft.py tests/test_gpu_api.py -x -q 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
