#!/bin/bash
# 8-GPU run: correctness at 8 ranks, collective sweeps vs NCCL, bench at N=8 and N=4
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout 900 python -m pytest tests/test_multirank.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_multirank8.log 2>&1
echo "pytest multirank exit $?" >> gpurun_out/pytest_multirank8.log
python -m mpi4jax_b200.run -n 8 --timeout 600 bench/collectives_sweep.py --out gpurun_out/collectives_sweep_n8.json > gpurun_out/sweep8.log 2>&1
echo "sweep exit $?" >> gpurun_out/sweep8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 1000 --warmup 50 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench8 exit $?" >> gpurun_out/bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 1000 --warmup 50 --no-sweep > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "bench4 exit $?" >> gpurun_out/bench_n4.err
python -m mpi4jax_b200.run -n 8 --timeout 300 scripts/gemm_perf.py > gpurun_out/gemm_perf_n8.log 2>&1
grep 'M=' gpurun_out/gemm_perf_n8.log
tail -n 3 gpurun_out/pytest_multirank8.log gpurun_out/sweep8.log gpurun_out/bench_n8.err gpurun_out/bench_n4.err
cut -c1-400 gpurun_out/bench_n8.json gpurun_out/bench_n4.json
