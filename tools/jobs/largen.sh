#!/bin/bash
python -m pytest tests/test_gpu_largen_hamsoft.py tests/test_gpu_largen.py -x -q 2>&1 | tail -12
python bench.py --workload largen --steps 3 --no-cpu 2>&1 | grep "^{" > gpurun_out/bench_largen.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_largen.json'))
print(d['value'], d['roofline']['frac'], d['checks'])
print(d['hamsoft_strang_substep'])
PY
