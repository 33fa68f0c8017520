"""Data-parallel training with differentiable allreduce (the README pattern of the reference:
average a loss / its gradients over the ranks).

    $ python -m mpi4jax_b200.run -n 4 examples/data_parallel_sgd.py [--cpu]

Every rank draws its own shard of a synthetic regression problem; parameters are broadcast from
rank 0, gradients are averaged with one fused allreduce per step, and all ranks end with the same
weights (checked at the end)."""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import mpi4jax_b200 as mpi4jax  # noqa: E402
from mpi4jax_b200 import MPI  # noqa: E402
from mpi4jax_b200.parallel import average_gradients, broadcast_parameters  # noqa: E402


def main(steps: int = 60, verbose: bool = True) -> float:
    comm = MPI.COMM_WORLD
    rank, size = comm.Get_rank(), comm.Get_size()
    dev = comm.device
    gen = torch.Generator().manual_seed(1234 + rank)
    w_true = torch.tensor([1.5, -2.0, 0.5], device=dev)
    x = torch.randn(256, 3, generator=gen).to(dev)
    y = x @ w_true + 0.01 * torch.randn(256, generator=gen).to(dev)

    w = torch.zeros(3, device=dev, requires_grad=True) if rank == 0 else \
        torch.full((3,), float(rank), device=dev, requires_grad=True)
    broadcast_parameters([w], root=0, comm=comm)             # everybody starts from rank 0's values
    opt = torch.optim.SGD([w], lr=0.1)
    loss_global = float("nan")
    for _ in range(steps):
        opt.zero_grad()
        loss = ((x @ w - y) ** 2).mean()
        loss.backward()
        average_gradients([w], comm=comm)                     # one allreduce(SUM) / size
        opt.step()
        loss_global = (mpi4jax.allreduce(loss.detach(), MPI.SUM, comm=comm) / size).item()
    # identical replicas: max - min over ranks is exactly zero
    spread = mpi4jax.allreduce(w.detach(), MPI.MAX, comm=comm) - mpi4jax.allreduce(w.detach(), MPI.MIN, comm=comm)
    assert float(spread.abs().max()) == 0.0
    mpi4jax.flush()
    if verbose and rank == 0:
        print(f"final loss {loss_global:.3e}, weights {w.detach().cpu().tolist()}")
    return loss_global


if __name__ == "__main__":
    main()
