#!/bin/bash
# ncu captures of hamsoft_run_kernel<3>: README systems (launch 3) and generic compact systems (launch 5)
python tools/profile_hamsoft.py 131072 100 2>&1 | tail -8 | tee gpurun_out/hs_speed.log
ncu --set full --clock-control none --import-source on --kernel-name regex:hamsoft_run_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r2_hs_n3 -f python tools/profile_hamsoft.py 32768 20 > gpurun_out/ncu_hs1.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:hamsoft_run_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/r2_hs_n3gen -f python tools/profile_hamsoft.py 32768 20 > gpurun_out/ncu_hs2.log 2>&1
tail -3 gpurun_out/ncu_hs2.log
