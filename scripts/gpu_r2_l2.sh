#!/bin/bash
# round 2, call L2 (2 GPUs): re-check after the allreduce_ validation and MLP capture fixes
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
timeout 200 python -m mpi4jax_b200.run -n 2 --timeout 180 --output-dir gpurun_out/r2l2_pytest_n2 -m pytest tests/test_extensions.py tests/test_models.py tests/test_coresidency.py \
   -q -m gpu -p no:cacheprovider -rf > /dev/null 2>&1
echo "pytest n2 exit $?"; tail -n 25 gpurun_out/r2l2_pytest_n2/rank0.log | cut -c1-250
timeout 120 python -m mpi4jax_b200.run -n 2 --timeout 100 bench/mlp_grad.py --out gpurun_out/r2l2_mlp_grad_n2.json > gpurun_out/r2l2_mlp_grad_n2.log 2>&1
grep -v "UserWarning\|run_backward\|^$" gpurun_out/r2l2_mlp_grad_n2.log | cut -c1-400 | tail -n 30
