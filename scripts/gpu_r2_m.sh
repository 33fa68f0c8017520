#!/bin/bash
# round 2, call M (8 GPUs): final multi-rank record -- GPU suite with per-test lines, the N=8 scaling point,
# full collectives sweep (staged, symmetric in-place, NCCL), phase timeline, MLP step, profiles with real peers
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
T0=$(date +%s)
stamp() { echo "== $1 (+$(( $(date +%s) - T0 )) s)"; }
stamp "pytest 8 ranks"
timeout 300 python -m mpi4jax_b200.run -n 8 --timeout 280 --output-dir gpurun_out/r2m_pytest_n8 -m pytest tests \
   -v -m gpu -p no:cacheprovider -rf > /dev/null 2>&1
echo "pytest n8 exit $?"; tail -n 3 gpurun_out/r2m_pytest_n8/rank0.log | cut -c1-200; grep -h "^FAILED\|^ERROR" gpurun_out/r2m_pytest_n8/rank*.log | sort | uniq -c | head -n 12
stamp "bench n=8 k=20"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 \
    bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2m_bench_n8_k20.json 2> gpurun_out/r2m_bench_n8_k20.err
python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/r2m_bench_n8_k20.json") if l.startswith("{")][-1]
    print({k: d.get(k) for k in ("n_gpus", "value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["checks"].get("checks_ok"), d.get("clocks"))
    print({k: {s: v["busbw"] for s, v in t.items()} for k, t in d["allreduce_busbw_gbs"].items()})
except Exception as e:
    print("bench parse error", e)
PY
tail -n 2 gpurun_out/r2m_bench_n8_k20.err | cut -c1-300
stamp "sweep"
timeout 400 python -m mpi4jax_b200.run -n 8 --timeout 380 bench/collectives_sweep.py --quick --skip-allreduce-algos \
   --out gpurun_out/r2m_sweep_n8.json > gpurun_out/r2m_sweep_n8.log 2>&1
echo "sweep exit $?"; grep -E "^fp32|^bf16" gpurun_out/r2m_sweep_n8.log | cut -c1-260; grep -E "^rooted|^allgather|^p2p" gpurun_out/r2m_sweep_n8.log | cut -c1-700
stamp "phases"
timeout 100 python -m mpi4jax_b200.run -n 8 --timeout 90 scripts/allreduce_phases.py 64 256 > gpurun_out/r2m_allreduce_phases_n8.log 2>&1
grep -v "^$" gpurun_out/r2m_allreduce_phases_n8.log | cut -c1-330 | tail -n 24
stamp "mlp grad"
timeout 100 python -m mpi4jax_b200.run -n 8 --timeout 90 bench/mlp_grad.py --out gpurun_out/r2m_mlp_grad_n8.json > gpurun_out/r2m_mlp_grad_n8.log 2>&1
grep -E "^dp_|^tp_" gpurun_out/r2m_mlp_grad_n8.log | cut -c1-300
stamp "ncu rank 0 (non-multicast kernels, full set)"
timeout 130 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29660 --no-python \
   scripts/rank0_ncu.sh x --set full --import-source on -k regex:'b2_k_move|b2_k_p2p|b2_k_halo_ca|swe_ca_|b2_k_reduce_chunked|b2_k_allreduce_ll' -s 10 -c 9 -o gpurun_out/r2m_rank0_n8 -- \
   scripts/prof_collectives.py > gpurun_out/r2m_ncu_n8.log 2>&1
echo "ncu rc=$?"; grep -c "==PROF== Profiling" gpurun_out/r2m_ncu_n8.log; tail -n 2 gpurun_out/r2m_ncu_n8.log | cut -c1-200
stamp "done"
