import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mpi4jax_b200 as m
from mpi4jax_b200 import MPI
from mpi4jax_b200.utils import max_over_ranks
comm = MPI.COMM_WORLD; r, n = comm.Get_rank(), comm.Get_size(); dev = comm.device
for nbytes in (1 << 20, 1 << 24, 1 << 28):
    x = torch.ones(nbytes // 4, device=dev)
    f = lambda: m.sendrecv(x, x, source=(r - 1) % n, dest=(r + 1) % n, comm=comm)
    for _ in range(3): f()
    torch.cuda.synchronize(); comm.Barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): f()
    e.record(); e.synchronize()
    us = max_over_ranks(s.elapsed_time(e) * 100, comm)
    if r == 0: print(f"sendrecv slot={os.environ.get('MPI4JAX_B200_P2P_SLOT_BYTES')} {nbytes}B {us:.1f}us {nbytes/us/1e3:.0f}GB/s", flush=True)
m.flush()
