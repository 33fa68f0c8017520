"""Distributed matrix transpose with alltoall (the pattern of the reference's regression test
tests/collective_ops/test_alltoall.py:43-65, and of Ulysses-style sequence <-> head resharding).

    $ python -m mpi4jax_b200.run -n 4 examples/distributed_transpose.py [--cpu]

A global (N x N) matrix is distributed by rows; after the transpose every rank owns the matching
row block of A^T.  The non-contiguous pack / unpack around the exchange is plain tensor
reshaping -- the op accepts non-contiguous inputs."""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import mpi4jax_b200 as mpi4jax  # noqa: E402
from mpi4jax_b200 import MPI  # noqa: E402


def transpose_rows(local: torch.Tensor, comm) -> torch.Tensor:
    """local: (n, N) row block of A  ->  (n, N) row block of A^T."""
    size = comm.Get_size()
    n, N = local.shape
    assert N == n * size
    blocks = local.reshape(n, size, n).permute(1, 0, 2)            # (size, n, n): block q goes to rank q
    recv = mpi4jax.alltoall(blocks, comm=comm)                     # recv[q] = block (q, rank) of A
    return recv.permute(2, 0, 1).reshape(n, N)                     # transpose every block, lay them out


def main(n: int = 3, verbose: bool = True) -> bool:
    comm = MPI.COMM_WORLD
    rank, size = comm.Get_rank(), comm.Get_size()
    N = n * size
    full = torch.arange(N * N, dtype=torch.float32, device=comm.device).reshape(N, N)
    mine = full[rank * n:(rank + 1) * n]
    out = mpi4jax.jit(lambda t: transpose_rows(t, comm))(mine)
    ok = torch.equal(out, full.T[rank * n:(rank + 1) * n])
    mpi4jax.flush()
    if verbose and rank == 0:
        print("distributed transpose", "ok" if ok else "WRONG")
    return ok


if __name__ == "__main__":
    sys.exit(0 if main() else 1)
