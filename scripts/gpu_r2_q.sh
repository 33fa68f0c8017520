#!/bin/bash
# round 2, call Q (2 GPUs): the final tree on two ranks -- whole GPU suite (fused GEMM + allreduce with the banded
# tile order included)
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
timeout 300 python -m mpi4jax_b200.run -n 2 --timeout 280 --output-dir gpurun_out/r2q_pytest_n2 -m pytest tests \
   -q -m gpu -p no:cacheprovider -rf > /dev/null 2>&1
echo "pytest n2 exit $?"; tail -n 6 gpurun_out/r2q_pytest_n2/rank0.log | cut -c1-250
timeout 100 python -m mpi4jax_b200.run -n 2 --timeout 90 scripts/gemm_perf.py 2>&1 | grep "^M=" | cut -c1-400
