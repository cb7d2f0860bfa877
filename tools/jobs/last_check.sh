#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests/test_gpu_classic.py tests/test_gpu_hamsoft.py tests/test_gpu_adaptive.py tests/test_gpu_largen.py -x -q 2>&1 | tail -2
python bench.py --no-largen --no-secondary --no-cpu --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('C3 %.3e (%.2f ms) e2e %.3e frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']), d['checks'])"
