#!/bin/bash
for v in A B C D; do
  echo "== $v"
  python tools/lib_override.py tools/variants/lib_$v.so --workload c1 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c1 value %.4e  e2e %.4e  ms %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
"
  python tools/hs_variant.py tools/variants/lib_$v.so 65536 100 2>&1 | tail -5
done
