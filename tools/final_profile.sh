# Round-end evidence run (1x B200): tests, default bench, launch list of one timed step, secondary workloads.
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; tail -3 gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
(time timeout 600 python bench.py) > gpurun_out/final_bench_n1.log 2>&1; tail -4 gpurun_out/final_bench_n1.log | cut -c1-400
(time timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/final_bench_ref.log 2>&1; tail -4 gpurun_out/final_bench_ref.log | cut -c1-300
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-largen > gpurun_out/final_bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 60 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-largen > gpurun_out/final_ncu.log 2>&1
tail -3 gpurun_out/final_launches.csv | cut -c1-200
timeout 600 python bench.py --workload largen > gpurun_out/final_bench_largen.log 2>&1; tail -1 gpurun_out/final_bench_largen.log | cut -c1-300
timeout 600 python bench.py --workload c4 > gpurun_out/final_bench_c4.log 2>&1; tail -1 gpurun_out/final_bench_c4.log | cut -c1-300
timeout 600 python bench.py --workload c1 > gpurun_out/final_bench_c1.log 2>&1; tail -1 gpurun_out/final_bench_c1.log | cut -c1-300
