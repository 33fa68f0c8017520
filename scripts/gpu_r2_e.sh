#!/bin/bash
# round 2, call E (4 GPUs): full multi-rank GPU suite, collectives sweep (all ops, NCCL lines), bench N=4, rank-0 launch profile
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=60
timeout 1200 python -m mpi4jax_b200.run -n 4 --timeout 1150 -m pytest tests/collective_ops tests/test_extensions.py tests/test_examples.py \
   tests/test_models.py tests/test_transport.py tests/test_jit.py tests/test_common.py tests/test_gemm.py tests/test_transforms.py \
   tests/test_compile.py tests/test_object_api.py tests/test_more_examples.py -q -m gpu -p no:cacheprovider > gpurun_out/r2e_pytest_n4.log 2>&1
echo "pytest n4 exit $?"; grep -E "passed|failed" gpurun_out/r2e_pytest_n4.log | tail -n 4; grep -E "^FAILED" gpurun_out/r2e_pytest_n4.log | sort | uniq | head -n 20
timeout 900 python -m mpi4jax_b200.run -n 4 --timeout 850 bench/collectives_sweep.py --quick --skip-allreduce-algos \
   --out gpurun_out/r2e_sweep_n4.json > gpurun_out/r2e_sweep_n4.log 2>&1
echo "sweep exit $?"; tail -n 30 gpurun_out/r2e_sweep_n4.log | cut -c1-400
for k in 200 20; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2951$((k % 7)) \
   bench.py --gpus 4 --steps $k --warmup 10 --no-sweep > gpurun_out/r2e_bench_n4_k$k.json 2> gpurun_out/r2e_bench_n4_k$k.err
echo "bench n4 k=$k rc=$?"; cut -c1-330 gpurun_out/r2e_bench_n4_k$k.json; tail -n 3 gpurun_out/r2e_bench_n4_k$k.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 --no-python \
   scripts/rank0_ncu.sh x --metrics gpu__time_duration.sum -s 40 -c 60 --csv --log-file gpurun_out/r2e_launches_n4_rank0.csv -- \
   scripts/swe_steps.py 4096 6 > gpurun_out/r2e_ncu_n4.log 2>&1
echo "ncu rc=$?"; tail -n 2 gpurun_out/r2e_ncu_n4.log
