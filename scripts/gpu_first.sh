#!/bin/bash
# first GPU bring-up: diagnostic + per-op GPU tests at world size 2 and 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/smi.txt 2>&1
N=${1:-2}
python -m mpi4jax_b200.run -n $N --timeout 500 scripts/gpu_diag.py > gpurun_out/diag.log 2>&1
echo "diag exit $?" >> gpurun_out/diag.log
python -m mpi4jax_b200.run -n $N --timeout 400 -m pytest tests/collective_ops -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_n$N.log 2>&1
echo "pytest n=$N exit $?" >> gpurun_out/pytest_n$N.log
timeout 300 python -m pytest tests/collective_ops -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_n1.log 2>&1
echo "pytest n=1 exit $?" >> gpurun_out/pytest_n1.log
tail -5 gpurun_out/diag.log gpurun_out/pytest_n$N.log gpurun_out/pytest_n1.log
