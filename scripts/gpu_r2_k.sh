#!/bin/bash
# round 2, call K (4 GPUs): new collective kernels (fused move loops, 2 CTAs/SM, pipelined reduce-to-root),
# phase timeline of the NVLS allreduce, the N=4 scaling point with the host-latency-free timed window
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
T0=$(date +%s)
stamp() { echo "== $1 (+$(( $(date +%s) - T0 )) s)"; }
stamp "pytest 4 ranks"
timeout 300 python -m mpi4jax_b200.run -n 4 --timeout 280 --output-dir gpurun_out/r2k_pytest_n4 -m pytest tests/collective_ops tests/test_extensions.py \
   tests/test_coresidency.py tests/test_models.py tests/test_examples.py tests/test_jit.py tests/test_gemm.py -q -m gpu -p no:cacheprovider -rf -x > /dev/null 2>&1
echo "pytest n4 exit $?"; tail -n 25 gpurun_out/r2k_pytest_n4/rank0.log | cut -c1-250
for n in 4 2 1; do
  stamp "bench n=$n k=20"
  if [ $n = 1 ]; then
    timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2k_bench_n1_k20.json 2> gpurun_out/r2k_bench_n1_k20.err
  else
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2k_bench_n${n}_k20.json 2> gpurun_out/r2k_bench_n${n}_k20.err
  fi
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/r2k_bench_n${n}_k20.json") if l.startswith("{")][-1]
    print({k: d.get(k) for k in ("n_gpus", "value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["checks"].get("checks_ok"), d.get("clocks"))
    if "allreduce_busbw_gbs" in d:
        print({k: {s: v["busbw"] for s, v in t.items()} for k, t in d["allreduce_busbw_gbs"].items()})
except Exception as e:
    print("bench n=${n} parse error", e)
PY
  tail -n 2 gpurun_out/r2k_bench_n${n}_k20.err | cut -c1-300
done
stamp "timeline n=4"
timeout 120 python -m mpi4jax_b200.run -n 4 --timeout 100 --output-dir gpurun_out/r2k_timeline_n4 scripts/swe_timeline.py 4096 6 > /dev/null 2>&1
tail -n 8 gpurun_out/r2k_timeline_n4/rank0.log
stamp "phases"
timeout 120 python -m mpi4jax_b200.run -n 4 --timeout 100 scripts/allreduce_phases.py 16 64 256 > gpurun_out/r2k_allreduce_phases_n4.log 2>&1
grep -v "^$" gpurun_out/r2k_allreduce_phases_n4.log | cut -c1-330 | tail -n 40
stamp "bw probe"
timeout 200 python -m mpi4jax_b200.run -n 4 --timeout 180 scripts/bw_probe.py 256 > gpurun_out/r2k_bw_probe_n4.log 2>&1
grep -E "MiB|nccl" gpurun_out/r2k_bw_probe_n4.log | cut -c1-400 || tail -n 5 gpurun_out/r2k_bw_probe_n4.log
stamp "sweep"
timeout 300 python -m mpi4jax_b200.run -n 4 --timeout 280 bench/collectives_sweep.py --quick --skip-allreduce-algos --min-bytes 1048576 \
   --out gpurun_out/r2k_sweep_n4.json > gpurun_out/r2k_sweep_n4.log 2>&1
echo "sweep exit $?"; grep -E "^fp32" gpurun_out/r2k_sweep_n4.log | cut -c1-200; grep -E "^rooted|^allgather|^p2p" gpurun_out/r2k_sweep_n4.log | cut -c1-600
stamp "mlp grad"
timeout 120 python -m mpi4jax_b200.run -n 4 --timeout 100 bench/mlp_grad.py --out gpurun_out/r2k_mlp_grad_n4.json > gpurun_out/r2k_mlp_grad_n4.log 2>&1
grep -E "^dp_|^tp_" gpurun_out/r2k_mlp_grad_n4.log | cut -c1-300
stamp "done"
