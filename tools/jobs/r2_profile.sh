#!/bin/bash
# Round-2 evidence run on one B200 (numbers are never taken under ncu; ncu runs come after the plain runs).
# gpurun copies back at most 64 MiB: every .ncu-rep is summarised on the box (tools/ncu_key.py, tools/ncu_lines.py) and
# only the two smallest reports travel.
O=gpurun_out
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_c4_horizon.py::test_million_steps_on_65536_systems 2>&1 | tail -2
( time python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err ) 2> $O/r2_bench_n1.time; tail -3 $O/r2_bench_n1.time | head -1
( time python bench.py --impl reference --steps 5 --warmup 1 > $O/r2_bench_ref.json 2>/dev/null ) 2> $O/r2_bench_ref.time
python bench.py --workload c1 > $O/r2_bench_c1.json 2>/dev/null
python bench.py --workload c4 > $O/r2_bench_c4.json 2>/dev/null
python bench.py --workload largen > $O/r2_bench_largen.json 2>/dev/null
python bench.py --workload c2 > $O/r2_bench_c2.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2_launches_bench_step.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-largen --no-secondary > $O/r2_ncu_launch.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name regex:ensemble_main_kernel -c 60 --csv --log-file $O/r2_main_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-largen --no-secondary > $O/r2_ncu_dram.log 2>&1
F="--set full --clock-control none --import-source on"
cap() {  # name, kernel regex, launch-skip, launch-count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu $F --kernel-name regex:$rx --launch-skip $skip --launch-count $cnt -o $O/$name -f "$@" > $O/${name}_ncu.log 2>&1
  python tools/ncu_key.py $O/$name.ncu-rep $O/${name}_key_metrics.csv > /dev/null 2>&1
  python tools/ncu_lines.py $O/$name.ncu-rep 40 > $O/${name}_hot_lines.txt 2>&1
  grep -h "TFLOP\|system-steps\|pairs/s\|pair evaluations" $O/${name}_ncu.log | tail -3
}
cap r2_main_N3 ensemble_main_kernel 2 1 python tools/profile_main.py 3
cap r2_main_N8 ensemble_main_kernel 2 1 python tools/profile_main.py 8
cap r2_hamsoft_N3 hamsoft_run_kernel 2 1 python tools/profile_hamsoft.py 131072 100
cap r2_whfast_N4 ensemble_main_kernel 2 1 python tools/profile_whfast.py 131072 100
cap r2_largen_x2 largeN_accel_x2 2 1 python tools/profile_largen.py 262144
cap r2_largen_pass largeN_pass_kernel 2 1 python tools/profile_pass.py 262144
rm -f $O/r2_main_N8.ncu-rep $O/r2_whfast_N4.ncu-rep $O/r2_largen_pass.ncu-rep $O/r2_largen_x2.ncu-rep
du -sh $O; ls $O | head -60
