#!/bin/bash
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=20
timeout 600 python -m pytest tests/test_examples.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_q1.log 2>&1
echo "exit $?" >> gpurun_out/pytest_q1.log
timeout 600 python bench.py --steps 400 --warmup 20 > gpurun_out/bench_q1.json 2> gpurun_out/bench_q1.err
MPI4JAX_B200_SWE_FUSED=0 timeout 600 python bench.py --steps 400 --warmup 20 > gpurun_out/bench_q1_unfused.json 2>> gpurun_out/bench_q1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 15 --csv --log-file gpurun_out/fused_launches.csv python scripts/swe_steps.py 4096 6 > gpurun_out/ncu_fused.log 2>&1
MPI4JAX_B200_SWE_FUSED=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 27 --csv --log-file gpurun_out/unfused_launches.csv python scripts/swe_steps.py 4096 6 > gpurun_out/ncu_unfused.log 2>&1
grep -E "AssertionError|passed|failed" gpurun_out/pytest_q1.log | head -8
cut -c1-260 gpurun_out/bench_q1.json gpurun_out/bench_q1_unfused.json
