"""reduce: the root receives the reduction over all ranks, every other rank gets its own input
back (reference: /root/reference/mpi4jax/_src/collective_ops/reduce.py:45-72, 124-133)."""

import pytest
import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI


comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()
ROOTS = sorted({0, size // 2, size - 1})


def _mine(device, dtype=torch.float32):
    return torch.arange(5, dtype=dtype, device=device) + rank


@pytest.mark.parametrize("root", ROOTS)
def test_sum_lands_on_the_root_only(device, root):
    x = _mine(device)
    keep = x.clone()
    out = m.reduce(x, op=MPI.SUM, root=root)
    if rank == root:
        want = torch.arange(5, dtype=torch.float32, device=device) * size + size * (size - 1) // 2
        assert torch.equal(out, want)
    else:
        assert torch.equal(out, x)                           # non-roots: the input comes back
    assert torch.equal(x, keep)


@pytest.mark.parametrize("name, dtype, expect", [
    ("MAX", torch.float32, lambda: torch.arange(5) + size - 1),
    ("MIN", torch.int64, lambda: torch.arange(5)),
    ("PROD", torch.float64, lambda: torch.stack([torch.arange(5.0, dtype=torch.float64) + r
                                                 for r in range(size)]).prod(0)),
], ids=["max", "min", "prod"])
def test_other_operators_on_the_last_rank(device, name, dtype, expect):
    root = size - 1
    x = _mine(device, dtype)
    out = m.reduce(x, op=getattr(MPI, name), root=root)
    if rank == root:
        assert torch.equal(out.cpu(), expect().to(dtype))
    else:
        assert torch.equal(out, x)


def test_python_scalar(device):
    out = m.reduce(rank, op=MPI.SUM, root=0)
    assert out.item() == (size * (size - 1) // 2 if rank == 0 else rank)


def test_replay_under_jit(device):
    x = torch.ones((3, 2), device=device) * rank
    total = m.jit(lambda t: m.reduce(t, op=MPI.SUM, root=0))
    for _ in range(3):
        out = total(x)
        want = torch.ones((3, 2), device=device) * (size * (size - 1) // 2) if rank == 0 else x
        assert torch.equal(out, want)
