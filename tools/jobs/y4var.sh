#!/bin/bash
F="--no-largen --no-secondary --no-cpu --steps 20 --warmup 3"
for v in nbodysimproject_b200/libnbody_b200.so tools/variants/lib_r6.so tools/variants/lib_r2.so nbodysimproject_b200/libnbody_b200.so tools/variants/lib_r6.so tools/variants/lib_r2.so; do
  echo "== $v"
  python tools/lib_override.py $v $F 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('ms/step %.2f  value %.3e  e2e %.2f ms  frac %.3f  main_only %.3f  %s' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], r['frac'], r['main_phase_only']['frac'], d['checks']))
"
done
