"""Phase timeline of the NVLS allreduce kernel with REAL peers: CTA 0 of every rank stamps %globaltimer
after each phase of each of its chunks (communicator option "trace_ptr", csrc/b2_collectives.cu), one
launch per message size.  A profiler cannot do this (ncu replays the kernel on one rank while the peers
move on); the stamps cost one barrier per phase on one CTA.
usage: python -m mpi4jax_b200.run -n 8 scripts/allreduce_phases.py [MiB ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mpi4jax_b200 as m  # noqa: E402
from mpi4jax_b200 import MPI  # noqa: E402

comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()
dev = comm.device
nc = comm._native_comm()
sizes = [int(a) for a in sys.argv[1:]] or [16, 64, 256]
CAP = 256
names = ["staged", "next staged", "W1 (all staged)", "ld_reduce + st", "W2 (all landed)", "copied out"]
for mib in sizes:
    x = torch.ones((mib << 20) // 4, device=dev)
    for pipe in (0, 1):
        nc.set_option("nvls_pipeline", pipe)
        for _ in range(3):
            m.allreduce(x, MPI.SUM, comm=comm)
        torch.cuda.synchronize()
        buf = torch.zeros(CAP, dtype=torch.int64, device=dev)
        nc.set_option("trace_ptr", buf.data_ptr())
        nc.set_option("trace_cap", CAP)
        comm.Barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m.barrier(comm=comm)
        s.record()
        m.allreduce(x, MPI.SUM, comm=comm)
        e.record()
        torch.cuda.synchronize()
        nc.set_option("trace_ptr", 0)
        nc.set_option("trace_cap", 0)
        t = buf.cpu().tolist()
        n = max(i for i, v in enumerate(t) if v) + 1 if any(t) else 0
        if rank == 0:
            print(f"--- {mib} MiB, {size} ranks, nvls_pipeline={pipe}: kernel {s.elapsed_time(e) * 1e3:.1f} us "
                  f"(event to event), grid cap {nc.get_option('max_blocks')}, CTA 0 of rank 0, {(n - 1) // 6} chunk(s)")
            t0 = t[0]
            for k in range((n - 1) // 6):
                seg = t[1 + 6 * k: 1 + 6 * (k + 1)]
                prev = t[6 * k]
                row = []
                for name, v in zip(names, seg):
                    row.append(f"{name} +{(v - prev) / 1e3:.1f}")
                    prev = v
                print(f"  chunk {k}: t={(t[6 * k] - t0) / 1e3:7.1f} us | " + " | ".join(row))
            print(f"  total {(t[n - 1] - t0) / 1e3:.1f} us", flush=True)
m.flush()
