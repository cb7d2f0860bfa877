#!/bin/bash
for O in 8,7,6,5,4,3 8,7,6,5,4,3 4,8,7,6,5,3; do
python bench.py --no-largen --no-secondary --no-cpu --order $O 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$O', round(d['ms_per_step'],2), round(d['roofline']['frac'],3), round(d['e2e']['ms_per_step'],2))"
done
