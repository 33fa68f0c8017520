"""A handful of collectives / p2p ops / shallow-water steps with REAL peers, for `ncu` on rank 0
(scripts/rank0_ncu.sh): one launch of every transport.  Prints nothing interesting."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mpi4jax_b200 as m  # noqa: E402
from mpi4jax_b200 import MPI  # noqa: E402
from mpi4jax_b200.models import ShallowWaterConfig, ShallowWaterModel  # noqa: E402

comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()
dev = comm.device
big = torch.ones(16 << 20, device=dev)                  # 64 MiB fp32
bigh = torch.ones(32 << 20, device=dev, dtype=torch.bfloat16)
small = torch.ones(4096, device=dev)
a2a = torch.ones(size, (16 << 20) // 4 // size, device=dev)
nxt, prv = (rank + 1) % size, (rank - 1) % size
for _ in range(2):                                      # first pass = warm-up (skipped by ncu -s)
    m.allreduce(small, MPI.SUM, comm=comm)              # LL
    m.allreduce(big, MPI.SUM, comm=comm)                # NVLS (two-shot without multicast)
    m.allreduce(bigh, MPI.SUM, comm=comm)
    m.allreduce(big, MPI.MAX, comm=comm)                # two-shot
    m.alltoall(a2a, comm=comm)
    m.allgather(small, comm=comm)
    m.bcast(big, size - 1, comm=comm)
    m.reduce(big, MPI.SUM, size - 1, comm=comm)
    m.scan(big, MPI.SUM, comm=comm)
    m.sendrecv(small, small, source=prv, dest=nxt, comm=comm)
    m.sendrecv(big, big, source=prv, dest=nxt, comm=comm)
    m.barrier(comm=comm)
    torch.cuda.synchronize()
model = ShallowWaterModel(ShallowWaterConfig.for_resolution(4096, 4096), comm=comm, device=dev)
model.multistep(3)
torch.cuda.synchronize()
print("done", rank, float(model.h.mean()))
m.flush()
