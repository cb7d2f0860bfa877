#!/bin/bash
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_c4_horizon.py::test_million_steps_on_65536_systems 2>&1 | tail -3
python bench.py --workload c1 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c1 value %.4e  e2e %.4e' % (d['value'], d['e2e']['value']), d['checks'])"
