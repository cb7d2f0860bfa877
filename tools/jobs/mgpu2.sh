#!/bin/bash
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --systems 131072 --steps 3 --particles 131072 > gpurun_out/bench_n2_small.json 2> gpurun_out/bench_n2_small.err; echo rc=$?; tail -c 1500 gpurun_out/bench_n2_small.err | grep -v "^W\|^$" | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2_small.json'))
print('value',d['value'],'e2e',d['e2e']['value'],d['checks'])
print(d['largen']['value'], d['largen']['checks'], d['largen']['hamsoft_strang_substep'])
print(d['c1']['value'], d['c4']['value'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/probe_host_bw.py 2>&1 | grep probe
