#!/bin/bash
# The GPU experiments that were prepared (written, model / host-emulation tested, default-off)
# when round 1 ran out of GPU budget -- one 1-GPU call, ~4 minutes.  Results land in gpurun_out/.
#   1. fused flux+tendency / friction pipelines (k12=1, 2) vs the stand-alone kernels
#   2. the same with -DB2_SWE_EXPLICIT_ROUNDING=1 (pipelines must then agree to the bit)
#   3. banded GEMM tile order (-DB2_GEMM_RASTER_GROUP=8)
#   4. CUDA variants of the tests that were added after the last GPU run
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=20
run_k12() {
  MPI4JAX_B200_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_examples.py tests/test_compile.py -q -m gpu \
      -p no:cacheprovider > gpurun_out/pytest_experimental_$1.log 2>&1
  echo "exit $?" >> gpurun_out/pytest_experimental_$1.log
  timeout 300 python scripts/k12_bench.py > gpurun_out/k12_bench_$1.log 2>&1
  tail -n 3 gpurun_out/pytest_experimental_$1.log; grep nx= gpurun_out/k12_bench_$1.log
}
echo "== default build"; run_k12 default
timeout 120 python scripts/gemm_perf.py > gpurun_out/gemm_default.log 2>&1; grep "M=" gpurun_out/gemm_default.log | cut -c1-260
echo "== explicit rounding + banded GEMM order"
MPI4JAX_B200_NVCC_FLAGS="-DB2_SWE_EXPLICIT_ROUNDING=1 -DB2_GEMM_RASTER_GROUP=8" python -m mpi4jax_b200._src.native.build > gpurun_out/rebuild.log 2>&1 || tail -n 5 gpurun_out/rebuild.log
export MPI4JAX_B200_NVCC_FLAGS="-DB2_SWE_EXPLICIT_ROUNDING=1 -DB2_GEMM_RASTER_GROUP=8"
run_k12 explicit
timeout 120 python scripts/gemm_perf.py > gpurun_out/gemm_raster8.log 2>&1; grep "M=" gpurun_out/gemm_raster8.log | cut -c1-260
timeout 300 python -m pytest tests/test_examples.py tests/test_gemm.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_explicit_build.log 2>&1
echo "exit $?" >> gpurun_out/pytest_explicit_build.log; tail -n 3 gpurun_out/pytest_explicit_build.log
