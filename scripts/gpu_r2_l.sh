#!/bin/bash
# round 2, call L (2 GPUs): symmetric in-place allreduce, traceback of the MLP capture failure
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
T0=$(date +%s)
stamp() { echo "== $1 (+$(( $(date +%s) - T0 )) s)"; }
stamp "pytest 2 ranks (extensions, models)"
timeout 200 python -m mpi4jax_b200.run -n 2 --timeout 180 --output-dir gpurun_out/r2l_pytest_n2 -m pytest tests/test_extensions.py tests/test_models.py tests/test_coresidency.py \
   -q -m gpu -p no:cacheprovider -rf -x > /dev/null 2>&1
echo "pytest n2 exit $?"; tail -n 25 gpurun_out/r2l_pytest_n2/rank0.log | cut -c1-250
stamp "mlp grad"
timeout 120 python -m mpi4jax_b200.run -n 2 --timeout 100 bench/mlp_grad.py --out gpurun_out/r2l_mlp_grad_n2.json > gpurun_out/r2l_mlp_grad_n2.log 2>&1
grep -v "UserWarning\|run_backward\|^$" gpurun_out/r2l_mlp_grad_n2.log | cut -c1-300 | tail -n 60
stamp "sweep allreduce"
timeout 300 python -m mpi4jax_b200.run -n 2 --timeout 280 bench/collectives_sweep.py --quick --skip-allreduce-algos --only-allreduce \
   --out gpurun_out/r2l_ar_n2.json > gpurun_out/r2l_ar_n2.log 2>&1
echo "sweep exit $?"; grep -E "^fp32|^bf16" gpurun_out/r2l_ar_n2.log | cut -c1-260
stamp "done"
