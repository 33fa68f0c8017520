#!/bin/bash
# first GPU call of round 2: validate and time the experimental fused flux+tendency path
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=20
# optional second pass: MPI4JAX_B200_NVCC_FLAGS="-DB2_SWE_EXPLICIT_ROUNDING=1" python -m mpi4jax_b200._src.native.build  (bit-identical pipelines)
MPI4JAX_B200_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_examples.py -q -m gpu -k "k12 or reproducible" -p no:cacheprovider > gpurun_out/pytest_k12.log 2>&1
echo "exit $?" >> gpurun_out/pytest_k12.log
timeout 300 python scripts/k12_bench.py 2>&1 | tee gpurun_out/k12_bench.log | grep nx=
tail -n 5 gpurun_out/pytest_k12.log
