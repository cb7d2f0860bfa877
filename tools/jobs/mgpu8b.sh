#!/bin/bash
# final 8-GPU line of the round-2 code (bench only; the host-bandwidth probe is in profiles/r2_probe_host_bw_n8.jsonl)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 > gpurun_out/r2_bench_n8.json 2> gpurun_out/bench_n8.err; echo rc=$?; grep -v "^W\|^\*\|^$\|OMP_NUM" gpurun_out/bench_n8.err | tail -5
python - <<'PY'
import json
for ln in open('gpurun_out/r2_bench_n8.json'):
    if ln.startswith('{'): d=json.loads(ln)
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'frac',d['roofline']['frac'])
print(d['checks'])
print('largen',d['largen']['value'],d['largen']['roofline']['frac'],d['largen']['checks'],d['largen']['hamsoft_strang_substep']['ms'])
print('c1',d['c1']['value'],'c4',d['c4']['value'])
PY
