#!/bin/bash
# single-GPU ncu captures for profiles/ (one kernel family per capture, --set full)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"swe_k|b2_k_halo_ll" -s 9 -c 7 -o gpurun_out/swe_step_full python scripts/swe_steps.py 4096 3 > gpurun_out/ncu_swe.log 2>&1
cat > /tmp/gemm_once.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from mpi4jax_b200 import MPI
from mpi4jax_b200.ops import linear_allreduce
x = torch.randn(4096, 4096, device="cuda").bfloat16(); w = torch.randn(4096, 4096, device="cuda").bfloat16()
for _ in range(3): y = linear_allreduce(x, w, comm=MPI.COMM_WORLD)
torch.cuda.synchronize(); print(y.float().abs().mean().item())
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:b2_k_gemm -s 2 -c 1 -o gpurun_out/gemm_full python /tmp/gemm_once.py > gpurun_out/ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"b2_k_reduce|b2_k_allreduce_ll|b2_k_move" -c 4 -o gpurun_out/coll_p1_full python -c "
import sys; sys.path.insert(0,'.')
import torch, mpi4jax_b200 as m
from mpi4jax_b200 import MPI
x = torch.ones(1<<26, device='cuda')
for _ in range(2): y = m.allreduce(x, MPI.SUM)
z = m.allgather(x)
torch.cuda.synchronize()
" > gpurun_out/ncu_coll.log 2>&1
ls -la gpurun_out/*.ncu-rep
