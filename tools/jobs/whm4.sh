#!/bin/bash
for v in tools/variants/lib_m6.so tools/variants/lib_m8.so tools/variants/lib_m5.so; do
  echo "== $v"
  python tools/lib_override.py $v --workload c4 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c4 value %.4e  e2e %.4e  ms %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
"
done
python tools/wh_variant.py tools/variants/lib_m5.so 2>&1 | tail -4
