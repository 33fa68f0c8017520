#!/bin/bash
# round 2, call D (2 GPUs): multi-rank GPU suite with real peers; N=1 schedules; bench N=1 / N=2 as the driver runs it
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
timeout 900 python -m mpi4jax_b200.run -n 2 --timeout 850 -m pytest tests/test_examples.py tests/test_extensions.py \
   tests/collective_ops tests/test_models.py tests/test_transport.py tests/test_jit.py tests/test_common.py -x -q -m gpu \
   -p no:cacheprovider > gpurun_out/r2d_pytest_n2.log 2>&1
echo "pytest n2 exit $?"; tail -n 6 gpurun_out/r2d_pytest_n2.log
for occ in 2 3; do
  MPI4JAX_B200_SWE_K12_OCC=$occ timeout 300 python scripts/swe_pipelines_bench.py 4096x4096 1024x2048 > gpurun_out/r2d_pipelines_occ$occ.log 2>&1
  echo "occ=$occ"; grep nx= gpurun_out/r2d_pipelines_occ$occ.log
done
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err
echo "bench n1 rc=$?"; cut -c1-400 gpurun_out/r2d_bench_n1.json; tail -n 3 gpurun_out/r2d_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
echo "bench n2 rc=$?"; cut -c1-400 gpurun_out/r2d_bench_n2.json; tail -n 5 gpurun_out/r2d_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus 2 --steps 20 --warmup 5 --no-sweep > gpurun_out/r2d_bench_n2_k20.json 2>> gpurun_out/r2d_bench_n2.err
cut -c1-300 gpurun_out/r2d_bench_n2_k20.json
