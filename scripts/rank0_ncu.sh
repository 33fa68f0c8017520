#!/bin/bash
# torchrun --no-python scripts/rank0_ncu.sh <ncu-output-stem> <ncu options ...> -- script.py args...
# Rank 0 runs under ncu (one pass per launch unless a full set is requested), the others plain:
# device times / counters of the kernels WITH real peers (NVLink traffic included).
stem=$1; shift
opts=()
while [ "$1" != "--" ]; do opts+=("$1"); shift; done
shift
if [ "${RANK:-0}" = "0" ]; then
  exec ncu "${opts[@]}" --clock-control none python "$@"
else
  exec python "$@"
fi
