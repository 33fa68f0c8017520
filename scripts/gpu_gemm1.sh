#!/bin/bash
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=20
timeout 240 python -m pytest tests/test_gemm.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gemm1.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gemm1.log
tail -25 gpurun_out/pytest_gemm1.log
timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/gemm_perf1.log
import torch, sys
sys.path.insert(0, ".")
from mpi4jax_b200 import MPI
from mpi4jax_b200.ops import linear_allreduce
comm = MPI.COMM_WORLD
for (M, N, K) in [(4096, 4096, 4096), (8192, 8192, 8192), (8192, 4096, 1024)]:
    x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
    for f, name in ((lambda: linear_allreduce(x, w, comm=comm), "tcgen05"), (lambda: x @ w.t(), "cublas")):
        for _ in range(3): f()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10): f()
        e.record(); e.synchronize()
        ms = s.elapsed_time(e) / 10
        print(f"{name} {M}x{N}x{K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
PY
