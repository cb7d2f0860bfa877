#!/bin/bash
# device-resident C3 step with 1 / 2 / 3 buffer+stream sets (steps in flight), two launch orders
F="--no-largen --no-secondary --no-cpu --steps 20 --warmup 3"
for v in "--sets 1" "--sets 2" "--sets 2" "--sets 2 --order 8,5,4,6,7,3" "--sets 3" "--sets 1"; do
  echo "== $v"
  python bench.py $F $v 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('ms/step %.2f  value %.3e  e2e %.2f ms  frac %.3f  main_only %.3f  same %s' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], r['frac'], r['main_phase_only']['frac'], d['checks']))
"
done
