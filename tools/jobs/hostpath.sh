#!/bin/bash
python -m pytest tests/test_gpu_host_path.py tests/test_gpu_api.py tests/test_gpu_classic.py -x -q 2>&1 | tail -15
python bench.py --no-largen --no-secondary --no-cpu > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; tail -c 600 gpurun_out/bench_e2e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_e2e.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['e2e']['ms_per_step'], d['checks'])
PY
