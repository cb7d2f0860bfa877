# Heavy-threshold sweep (one B200): the automatic rule max(4, 50 / speedup(N)) against fixed thresholds.
# The threshold is an argument of nb_sort_by_nsub (DeviceBucket.sort(heavy_threshold=...)); bench.py --heavy-threshold T
# passes it through on the device-resident path.
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"]/1e9,3), round(d["ms_per_step"],1), "roof", round(r["frac"],3), round(r["ms"],1))'
for T in -1 4 8 16 32 63; do
  echo "heavy threshold = $T (-1 = automatic)"
  python bench.py --no-cpu --no-largen --no-secondary --heavy-threshold $T | grep "^{" | python -c "$P"
done
