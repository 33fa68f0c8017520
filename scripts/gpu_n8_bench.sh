#!/bin/bash
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 2000 --warmup 50 --no-sweep > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench8 exit $?" >> gpurun_out/bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 2000 --warmup 50 --no-sweep > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "bench4 exit $?" >> gpurun_out/bench_n4.err
cut -c1-330 gpurun_out/bench_n8.json gpurun_out/bench_n4.json
tail -n 1 gpurun_out/bench_n8.err gpurun_out/bench_n4.err
