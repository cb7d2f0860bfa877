#!/bin/bash
python bench.py --no-largen --no-secondary --no-cpu 2>gpurun_out/bench_c3.err | grep "^{" > gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3.json'))
print('c3 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['e2e']['ms_per_step'], d['checks'])
r=d['roofline']; print('roof',r['achieved'],r['frac'],r['frac_nominal'],r['ms'],r['share_of_step'])
for p in r['per_bucket']: print(p)
PY
