#!/bin/bash
# usage: tools/grun.sh <out-file> <timeout-s> [--gpus N] -- <command...> : gpurun with retries while the pod answers "transient"/busy
out=$1; shift; tmo=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $tmo "${extra[@]}" -- "$@" > "$out" 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes" "$out" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -3 "$out"
