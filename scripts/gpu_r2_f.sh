#!/bin/bash
# round 2, call F (8 GPUs): multi-rank GPU suite with per-rank logs, scaling points, timeline, sweeps, profiles
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
T0=$(date +%s)
stamp() { echo "== $1 (+$(( $(date +%s) - T0 )) s)"; }
stamp "pytest 8 ranks"
timeout 300 python -m mpi4jax_b200.run -n 8 --timeout 280 --output-dir gpurun_out/r2f_pytest_n8 -m pytest tests \
   -v -m gpu -p no:cacheprovider -rf > /dev/null 2>&1
echo "pytest n8 exit $?"; tail -n 3 gpurun_out/r2f_pytest_n8/rank0.log | cut -c1-200; grep -h "^FAILED\|^ERROR" gpurun_out/r2f_pytest_n8/rank*.log | sort | uniq -c | head -n 12
for n in 8 4; do
  stamp "bench n=$n k=20"
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2f_bench_n${n}_k20.json 2> gpurun_out/r2f_bench_n${n}_k20.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/r2f_bench_n${n}_k20.json") if l.startswith("{")][-1]
    print({k: d.get(k) for k in ("n_gpus", "value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["checks"].get("checks_ok"),
          {k: v for k, v in d["checks"].items() if v not in (0.0, True)}, d.get("public_ops_shallow_water"))
    if "allreduce_busbw_gbs" in d:
        print({k: {s: v["busbw"] for s, v in t.items()} for k, t in d["allreduce_busbw_gbs"].items()})
except Exception as e:
    print("bench n=${n} parse error", e)
PY
  tail -n 2 gpurun_out/r2f_bench_n${n}_k20.err | cut -c1-300
done
stamp "timeline"
timeout 120 python -m mpi4jax_b200.run -n 8 --timeout 100 --output-dir gpurun_out/r2f_timeline_n8 scripts/swe_timeline.py 4096 8 > /dev/null 2>&1
tail -n 10 gpurun_out/r2f_timeline_n8/rank0.log
stamp "sweep"
timeout 400 python -m mpi4jax_b200.run -n 8 --timeout 380 bench/collectives_sweep.py --quick --skip-allreduce-algos \
   --out gpurun_out/r2f_sweep_n8.json > gpurun_out/r2f_sweep_n8.log 2>&1
echo "sweep exit $?"; grep -E "^fp32|^bf16" gpurun_out/r2f_sweep_n8.log | cut -c1-200; grep -E "^rooted|^allgather|^p2p" gpurun_out/r2f_sweep_n8.log | cut -c1-700
stamp "bw probe"
timeout 200 python -m mpi4jax_b200.run -n 8 --timeout 180 scripts/bw_probe.py 64 > gpurun_out/r2f_bw_probe_n8.log 2>&1
grep -E "MiB|nccl" gpurun_out/r2f_bw_probe_n8.log | cut -c1-400 || tail -n 5 gpurun_out/r2f_bw_probe_n8.log
stamp "mlp grad"
timeout 120 python -m mpi4jax_b200.run -n 8 --timeout 100 bench/mlp_grad.py --out gpurun_out/r2f_mlp_grad_n8.json > gpurun_out/r2f_mlp_grad_n8.log 2>&1
tail -n 4 gpurun_out/r2f_mlp_grad_n8.log | cut -c1-300
stamp "ncu rank 0"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29660 --no-python \
   scripts/rank0_ncu.sh x --set full --import-source on -k regex:'b2_k_|swe_ca_' -s 14 -c 24 -o gpurun_out/r2f_rank0_n8 -- \
   scripts/prof_collectives.py > gpurun_out/r2f_ncu_n8.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/r2f_ncu_n8.log | cut -c1-200
stamp "done"
