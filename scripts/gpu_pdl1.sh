#!/bin/bash
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=20
timeout 300 python scripts/pdl_repro.py 2>&1 | tee gpurun_out/pdl_repro.log | grep -c IDENTICAL
grep -v IDENTICAL gpurun_out/pdl_repro.log | head -20
timeout 300 python scripts/pdl_bench.py 2>&1 | tee gpurun_out/pdl_bench.log | grep nx=
timeout 900 python -m pytest tests/test_examples.py tests/test_jit.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_new.log 2>&1
echo "exit $?" >> gpurun_out/pytest_new.log
tail -n 6 gpurun_out/pytest_new.log
