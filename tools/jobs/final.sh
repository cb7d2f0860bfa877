#!/bin/bash
# last run of the round: full GPU suite, smoke, default bench line, C4 line (final code)
mkdir -p gpurun_out
( time stdbuf -oL python -m pytest tests -m gpu -x -v --durations=10 ) 2>&1 | stdbuf -oL grep -v "^$" > gpurun_out/r2_pytest_gpu.log
tail -16 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo bench rc=$?
python bench.py --workload c4 > gpurun_out/r2_bench_c4.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('C3 %.4e (%.2f ms) e2e %.4e (%.2f ms) frac %.3f main_only %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], r['frac'], r['main_phase_only']['frac']))
print('largen %.4e c1 %.4e c4 %.4e' % (d['largen']['value'], d['c1']['value'], d['c4']['value']), d['checks'], d['clocks'])
PY
