"""Conjugate gradients with a column-sharded operator: every mat-vec is a local product plus an
allreduce (the reference's integration test, tests/test_jax_transforms.py:6-22, runs
jax.scipy.sparse.linalg.cg on such an operator).

    $ python -m mpi4jax_b200.run -n 4 examples/conjugate_gradient.py [--cpu]

The loop runs inside mpi4jax_b200.jit: on GPUs the whole fixed-iteration solve is one CUDA graph."""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import mpi4jax_b200 as mpi4jax  # noqa: E402
from mpi4jax_b200 import MPI  # noqa: E402


def main(n: int = 48, iters: int = 60, verbose: bool = True) -> float:
    comm = MPI.COMM_WORLD
    rank, size = comm.Get_rank(), comm.Get_size()
    dev = comm.device
    gen = torch.Generator().manual_seed(7)
    m = torch.randn(n, n, generator=gen, dtype=torch.float64)
    a = (m @ m.T + n * torch.eye(n, dtype=torch.float64)).to(dev)        # SPD, the same on every rank
    b = torch.randn(n, generator=gen, dtype=torch.float64).to(dev)
    cols = torch.arange(n, device=dev).chunk(size)[rank]                    # my column shard
    a_cols = a[:, cols]

    def matvec(x):                                # A @ x = sum over ranks of A[:, cols] @ x[cols]
        return mpi4jax.allreduce(a_cols @ x[cols], MPI.SUM, comm=comm)

    def solve(rhs):
        x = torch.zeros_like(rhs)
        r = rhs - matvec(x)
        p = r.clone()
        rs = r @ r
        for _ in range(iters):
            ap = matvec(p)
            alpha = rs / (p @ ap)
            x = x + alpha * p
            r = r - alpha * ap
            rs_new = r @ r
            p = r + (rs_new / rs) * p
            rs = rs_new
        return x

    x = mpi4jax.jit(solve)(b)
    x = mpi4jax.jit(solve)(b)                      # second call: captured / replayed on GPUs
    err = float((a @ x - b).norm() / b.norm())
    mpi4jax.flush()
    if verbose and rank == 0:
        print(f"relative residual after {iters} iterations: {err:.2e}")
    return err


if __name__ == "__main__":
    main()
