P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"]/1e9,3), round(d["ms_per_step"],1), round(d["e2e"]["value"]/1e9,3), "roof", round(r["frac"],3), round(r["ms"],1))'
for S in 0.5 0.75 1 1.5 2.5; do echo "speedup_scale=$S"; NB_HEAVY_SPEEDUP_SCALE=$S timeout 300 python bench.py --no-cpu --no-largen 2>/dev/null | grep "^{" | python -c "$P"; done
