#!/bin/bash
python bench.py --workload c4 --no-cpu --steps 3 2>&1 | grep "^{" > gpurun_out/bench_c4.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c4.json'))
print('c4', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['counted_work'], d['checks'])
PY
python -m pytest tests/test_gpu_classic.py tests/test_gpu_midn.py -x -q -k "whfast or kepler" 2>&1 | tail -3
