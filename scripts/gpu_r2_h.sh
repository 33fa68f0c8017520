#!/bin/bash
# round 2, call H (2 GPUs): shake-out before the 8-GPU call -- full GPU suite on 2 ranks, bench, timeline, quick sweep
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
T0=$(date +%s)
stamp() { echo "== $1 (+$(( $(date +%s) - T0 )) s)"; }
stamp "pytest 1 rank (examples, models)"
timeout 300 python -m pytest tests/test_examples.py tests/test_models.py -q -m gpu -p no:cacheprovider > gpurun_out/r2h_pytest_n1.log 2>&1
echo "pytest n1 exit $?"; tail -n 3 gpurun_out/r2h_pytest_n1.log | cut -c1-300
stamp "pytest 2 ranks"
timeout 420 python -m mpi4jax_b200.run -n 2 --timeout 400 --output-dir gpurun_out/r2h_pytest_n2 -m pytest tests \
   -q -m gpu -p no:cacheprovider -rf > /dev/null 2>&1
echo "pytest n2 exit $?"; tail -n 3 gpurun_out/r2h_pytest_n2/rank0.log | cut -c1-300; grep -h "^FAILED\|^ERROR" gpurun_out/r2h_pytest_n2/rank*.log | sort | uniq -c | head -n 12
stamp "bench n=2 k=20"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n2_k20.json 2> gpurun_out/r2h_bench_n2_k20.err
cut -c1-400 gpurun_out/r2h_bench_n2_k20.json; tail -n 2 gpurun_out/r2h_bench_n2_k20.err | cut -c1-300
stamp "timeline n=2"
timeout 200 python -m mpi4jax_b200.run -n 2 --timeout 180 --output-dir gpurun_out/r2h_timeline_n2 scripts/swe_timeline.py 4096 6 > /dev/null 2>&1
tail -n 8 gpurun_out/r2h_timeline_n2/rank0.log
stamp "sweep"
timeout 400 python -m mpi4jax_b200.run -n 2 --timeout 380 bench/collectives_sweep.py --quick --skip-allreduce-algos \
   --out gpurun_out/r2h_sweep_n2.json > gpurun_out/r2h_sweep_n2.log 2>&1
echo "sweep exit $?"; grep -E "^fp32|^bf16" gpurun_out/r2h_sweep_n2.log | cut -c1-200; grep -E "^rooted|^allgather|^p2p" gpurun_out/r2h_sweep_n2.log | cut -c1-700
stamp "done"
