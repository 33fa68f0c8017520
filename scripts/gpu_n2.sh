#!/bin/bash
# 2-GPU validation of the vectorised stencils + single-phase halo, perf snapshot
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=20
timeout 600 python -m pytest tests/test_examples.py tests/test_models.py tests/test_jit.py tests/test_common.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_a_n1.log 2>&1
echo "pytest n=1 exit $?" >> gpurun_out/pytest_a_n1.log
timeout 900 python -m pytest tests/test_multirank.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_multirank.log 2>&1
echo "pytest multirank exit $?" >> gpurun_out/pytest_multirank.log
python -m mpi4jax_b200.run -n 2 --timeout 500 scripts/gpu_diag.py > gpurun_out/diag.log 2>&1
echo "diag exit $?" >> gpurun_out/diag.log
timeout 600 python bench.py --steps 400 --warmup 20 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench1 exit $?" >> gpurun_out/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 400 --warmup 20 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench2 exit $?" >> gpurun_out/bench_n2.err
tail -n 4 gpurun_out/pytest_a_n1.log gpurun_out/pytest_multirank.log gpurun_out/bench_n1.err gpurun_out/bench_n2.err
grep -E "halo|swe|perf|FAIL" gpurun_out/diag.log | tail -12
cat gpurun_out/bench_n1.json gpurun_out/bench_n2.json | cut -c1-600
python scripts/swe_small_bench.py 2>&1 | tee gpurun_out/swe_small.log | grep nx=
