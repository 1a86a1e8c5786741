#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing charged).  usage: scripts/gpurun_retry.sh [--gpus N] <timeout_s> '<command>'
extra=""
if [ "$1" == "--gpus" ]; then extra="--gpus $2"; shift 2; fi
t=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun $extra --timeout $t -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy"; then echo "[retry $i] transient/busy"; sleep 90; continue; fi
  echo "$out" | tail -60
  exit 0
done
echo "gave up"
