P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"]/1e9,3), round(d["ms_per_step"],1), round(d["e2e"]["value"]/1e9,3), "roof", round(r["frac"],3), round(r["ms"],1), [(b["N"], round(b["solo_ms"],1)) for b in r["per_bucket_solo"]], d["checks"])'
for K in 1 2 4; do echo "kappa=$K"; NB_HEAVY_KAPPA=$K timeout 300 python bench.py --no-cpu --no-largen 2>gpurun_out/err_k$K.log | tail -1 | python -c "$P"; done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/check_thr.py 2>&1 | grep -v Warn
