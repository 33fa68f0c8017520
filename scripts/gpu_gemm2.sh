#!/bin/bash
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=20
N=${1:-2}
python -m mpi4jax_b200.run -n $N --timeout 300 -m pytest tests/test_gemm.py tests/test_models.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gemm_n$N.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gemm_n$N.log
tail -6 gpurun_out/pytest_gemm_n$N.log
python -m mpi4jax_b200.run -n $N --timeout 300 scripts/gemm_perf.py 2>&1 | tee gpurun_out/gemm_perf_n$N.log | grep "M="
