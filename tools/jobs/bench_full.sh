#!/bin/bash
( time python bench.py --no-largen-hamsoft > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2> gpurun_out/bench_full.time; echo rc=$?; tail -c 1500 gpurun_out/bench_full.err; cat gpurun_out/bench_full.time
