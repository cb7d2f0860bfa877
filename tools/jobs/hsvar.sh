#!/bin/bash
for v in nbodysimproject_b200/libnbody_b200.so tools/variants/lib_occ.so; do
  echo "== $v"
  python tools/lib_override.py $v --workload c1 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c1 value %.4e  e2e %.4e  ms %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
"
  python tools/hs_variant.py $v 65536 100 2>&1 | tail -5
done
