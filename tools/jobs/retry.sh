#!/bin/bash
# usage: tools/jobs/retry.sh <timeout> <gpus> <job script>; retries while the pod answers busy (nothing is charged then)
T=$1; G=$2; J=$3
for k in $(seq 1 20); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "bash $J" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $J" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient\|rc=3\|busy"; then sleep 45; continue; fi
  echo "$OUT"; exit 0
done
echo "$OUT"; echo "gave up after 20 tries"
