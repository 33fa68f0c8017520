#!/bin/bash
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=20
timeout 600 python -m pytest tests/test_examples.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_q1.log 2>&1
echo "exit $?" >> gpurun_out/pytest_q1.log
python scripts/swe_small_bench.py 2>&1 | tee gpurun_out/swe_small.log | grep nx=
grep -E "AssertionError|passed|failed" gpurun_out/pytest_q1.log | head -8
