#!/bin/bash
# round 2, call G5 (1 GPU): kernel A with plain fluxes at owned cells + called generic path
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
timeout 600 python -m pytest tests/test_examples.py tests/test_models.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2g5_pytest.log 2>&1
echo "pytest exit $?"; tail -n 4 gpurun_out/r2g5_pytest.log | cut -c1-300
timeout 200 python scripts/swe_timeline.py 4096 6 > gpurun_out/r2g5_timeline_4096.log 2>&1; tail -n 8 gpurun_out/r2g5_timeline_4096.log
timeout 200 python scripts/swe_timeline.py 1448 6 > gpurun_out/r2g5_timeline_1448.log 2>&1; tail -n 8 gpurun_out/r2g5_timeline_1448.log
timeout 300 python scripts/swe_pipelines_bench.py 4096x4096 1448x1448 > gpurun_out/r2g5_pipelines.log 2>&1
grep nx= gpurun_out/r2g5_pipelines.log || tail -n 20 gpurun_out/r2g5_pipelines.log
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/r2g5_bench_n1.json 2> gpurun_out/r2g5_bench_n1.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r2g5_bench_n1.json; tail -n 3 gpurun_out/r2g5_bench_n1.err
