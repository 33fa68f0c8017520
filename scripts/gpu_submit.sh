#!/bin/bash
# usage: scripts/gpu_submit.sh <tag> [gpurun options] -- <command>
# gpurun snapshots /root/repo when a box is granted, possibly minutes after the call: run the command
# from a frozen copy (.frozen/<tag>/, shipped with the snapshot) so that work can go on in the tree.
# Results still land in gpurun_out/ (symlinked from the copy).  Retries while the pod is busy.
set -e
tag=$1; shift
opts=()
while [ "$1" != "--" ]; do opts+=("$1"); shift; done
shift
root=$(cd "$(dirname "$0")/.." && pwd)
dst=$root/.frozen/$tag
rm -rf "$dst"; mkdir -p "$dst"
tar -C "$root" --exclude=./.git --exclude=./gpurun_out --exclude=./.frozen --exclude=./mpi4jax_b200/_native/obj \
    --exclude=__pycache__ --exclude=./.pytest_cache --exclude=./baseline/_ref -cf - . | tar -C "$dst" -xf -
ln -s ../../gpurun_out "$dst/gpurun_out"
mkdir -p "$root/gpurun_out"
log=$root/gpurun_out/${tag}_call.log
for i in $(seq 1 60); do
  set +e
  /usr/local/graft/bin/gpurun "${opts[@]}" -- "cd .frozen/$tag && $*" > "$log" 2>&1
  rc=$?
  set -e
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then echo "done rc=$rc" >> "$log"; rm -rf "$dst"; exit $rc; fi
  sleep 30
done
echo "gave up" >> "$log"
