# Heavy-threshold sweep (one B200): the automatic rule max(4, 50 / speedup(N)) against fixed thresholds.
# nb_ensemble_set_heavy_nsub() is process-wide, so every point is its own process.
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"]/1e9,3), round(d["ms_per_step"],1), round(d["e2e"]["value"]/1e9,3), "roof", round(r["frac"],3), round(r["ms"],1))'
for T in -1 4 8 16 32 63; do
  echo "heavy threshold = $T (-1 = automatic)"
  python - "$T" <<'PY' | grep "^{" | python -c "$P"
import sys, runpy
sys.path.insert(0, ".")
from nbodysimproject_b200 import _lib as L
L.check(L.load().nb_ensemble_set_heavy_nsub(int(sys.argv[1])))
sys.argv = ["bench.py", "--no-cpu", "--no-largen"]
runpy.run_path("bench.py", run_name="__main__")
PY
done
