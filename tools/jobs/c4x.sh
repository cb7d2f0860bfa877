#!/bin/bash
( time python -m pytest tests/test_gpu_c4_horizon.py tests/test_gpu_classic.py -x -q -k "exact or whfast" 2>&1 | tail -40 ) 2>&1
