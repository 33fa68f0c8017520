#!/bin/bash
# round 2, call J (2 GPUs): full suite after the p2p / func.grad fixes, p2p latency, bandwidth probe
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
T0=$(date +%s)
stamp() { echo "== $1 (+$(( $(date +%s) - T0 )) s)"; }
stamp "pytest 2 ranks"
timeout 420 python -m mpi4jax_b200.run -n 2 --timeout 400 --output-dir gpurun_out/r2j_pytest_n2 -m pytest tests \
   -q -m gpu -p no:cacheprovider -rf -x > /dev/null 2>&1
echo "pytest n2 exit $?"; tail -n 30 gpurun_out/r2j_pytest_n2/rank0.log | cut -c1-250
stamp "p2p"
timeout 300 python -m mpi4jax_b200.run -n 2 --timeout 280 bench/collectives_sweep.py --quick --skip-allreduce-algos --only-p2p \
   --out gpurun_out/r2j_p2p_n2.json > gpurun_out/r2j_p2p_n2.log 2>&1
grep -E "^p2p" gpurun_out/r2j_p2p_n2.log | cut -c1-700 || tail -n 5 gpurun_out/r2j_p2p_n2.log
stamp "bw probe"
timeout 300 python -m mpi4jax_b200.run -n 2 --timeout 280 scripts/bw_probe.py 256 > gpurun_out/r2j_bw_probe_n2.log 2>&1
grep -E "MiB|nccl" gpurun_out/r2j_bw_probe_n2.log | cut -c1-400 || tail -n 5 gpurun_out/r2j_bw_probe_n2.log
stamp "pipelines 1 gpu (in-place layout)"
timeout 300 python scripts/swe_pipelines_bench.py 4096x4096 1448x1448 2048x1024 > gpurun_out/r2j_pipelines.log 2>&1
grep nx= gpurun_out/r2j_pipelines.log || tail -n 20 gpurun_out/r2j_pipelines.log
timeout 200 python scripts/swe_timeline.py 1448 6 > gpurun_out/r2j_timeline_1448.log 2>&1; tail -n 4 gpurun_out/r2j_timeline_1448.log
stamp "done"
