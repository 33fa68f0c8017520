"""alltoall: block q of rank r goes to block r of rank q (scenario parity with
/root/reference/tests/collective_ops/test_alltoall.py incl. the non-contiguous regression mpi4jax#176)."""

import pytest
import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

comm = MPI.COMM_WORLD
rank = comm.Get_rank()
size = comm.Get_size()


def test_alltoall(device):
    arr = torch.ones((size, 3, 2), device=device) * rank
    _arr = arr.clone()
    res = m.alltoall(arr)
    for p in range(size):
        assert torch.equal(res[p], torch.ones((3, 2), device=device) * p)
    assert torch.equal(_arr, arr)


def test_alltoall_jit(device):
    arr = torch.ones((size, 3, 2), device=device) * rank
    f = m.jit(lambda x: m.alltoall(x))
    for _ in range(3):
        res = f(arr)
        for p in range(size):
            assert torch.equal(res[p], torch.ones((3, 2), device=device) * p)


def test_alltoall_wrong_size(device):
    arr = torch.ones((size + 1, 3, 2), device=device) * rank
    with pytest.raises(ValueError) as excinfo:
        m.alltoall(arr)
    assert "must have shape (nproc, ...)" in str(excinfo.value)


def test_alltoall_transpose(device):
    """Distributed transpose through non-contiguous views (reference regression mpi4jax#176,
    test_alltoall.py:43-65)."""
    n = 4
    full = torch.arange(size * n * size * n, dtype=torch.float32).reshape(size * n, size * n)
    mine = full[rank * n:(rank + 1) * n].to(device)           # my row block (n, size*n)

    def dist_transpose(a):
        a = a.reshape(n, size, n).permute(1, 0, 2)            # (size, n, n): block q goes to rank q
        a = m.alltoall(a)                                     # handles the non-contiguous view
        return a.permute(2, 0, 1).reshape(n, size * n)        # rows of the transposed matrix

    for f in (dist_transpose, m.jit(dist_transpose)):
        for _ in range(3):
            res = f(mine)
            assert torch.equal(res.cpu(), full.t()[rank * n:(rank + 1) * n])


def test_alltoall_grad(device):
    x = (torch.ones((size, 3), device=device) * (rank + 1)).requires_grad_(True)
    out = m.alltoall(x)
    (out * (rank + 1)).sum().backward()
    exp = torch.stack([torch.ones(3) * (p + 1) for p in range(size)]).to(device)
    assert torch.equal(x.grad, exp)
