#!/bin/bash
# round 2, call A (1 GPU): explicit-rounding build; every gated scenario; K12 / K12f vs stand-alone
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=20
export MPI4JAX_B200_TEST_EXPERIMENTAL=1
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -rs > gpurun_out/r2a_pytest.log 2>&1
echo "exit $?" >> gpurun_out/r2a_pytest.log
tail -n 15 gpurun_out/r2a_pytest.log
timeout 300 python scripts/k12_bench.py > gpurun_out/r2a_k12_bench.log 2>&1
grep nx= gpurun_out/r2a_k12_bench.log || tail -n 20 gpurun_out/r2a_k12_bench.log
