#!/bin/bash
python bench.py --systems 65536 --steps 3 --particles 65536 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo rc=$?; tail -c 3000 gpurun_out/bench_small.err; head -c 6000 gpurun_out/bench_small.json
