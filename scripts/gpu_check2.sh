#!/bin/bash
# round-end style check on a 2-GPU box: full GPU suite, smoke, bench N=1 and N=2
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/final_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/final_smoke.log
timeout 600 python bench.py --gpus 1 --steps 2000 --warmup 50 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench1 exit $?" >> gpurun_out/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2000 --warmup 50 --no-sweep > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench2 exit $?" >> gpurun_out/bench_n2.err
tail -n 3 gpurun_out/final_pytest_gpu.log gpurun_out/final_smoke.log
cut -c1-400 gpurun_out/bench_n1.json gpurun_out/bench_n2.json
tail -n 2 gpurun_out/bench_n1.err gpurun_out/bench_n2.err
