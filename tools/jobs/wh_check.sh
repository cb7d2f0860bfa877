#!/bin/bash
python -m pytest tests/test_gpu_classic.py tests/test_gpu_midn.py tests/test_gpu_hamsoft_mid.py tests/test_gpu_c4_horizon.py::test_one_long_launch_equals_many_short_ones -x -q 2>&1 | tail -4
python tools/wh_variant.py nbodysimproject_b200/libnbody_b200.so 2>&1 | tail -4
python bench.py --workload c4 --no-cpu 2>/dev/null | tail -1 > gpurun_out/r2_bench_c4_v2.json
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_c4_v2.json').read())
print('c4 value %.3e e2e %.3e frac %.4f' % (d['value'], d['e2e']['value'], d['roofline']['frac']))"
