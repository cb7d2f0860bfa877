#!/bin/bash
for v in nbodysimproject_b200/libnbody_b200.so tools/variants/lib_base.so tools/variants/lib_A.so tools/variants/lib_B.so tools/variants/lib_C.so; do
  echo "== $v"; python tools/wh_variant.py $v 2>&1 | tail -4
done
