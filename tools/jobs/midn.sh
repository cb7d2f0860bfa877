#!/bin/bash
python -m pytest tests/test_gpu_midn.py tests/test_gpu_classifier.py tests/test_gpu_host_path.py tests/test_gpu_largen.py tests/test_gpu_classic.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -25
