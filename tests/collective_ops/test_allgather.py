"""allgather: every rank receives every rank's array, `S -> (nproc, *S)` (scenario parity with
/root/reference/tests/collective_ops/test_allgather.py; plus dtypes, autograd and large messages)."""

import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

comm = MPI.COMM_WORLD
rank = comm.Get_rank()
size = comm.Get_size()


def test_allgather(device):
    arr = torch.ones((3, 2), device=device) * rank
    _arr = arr.clone()
    res = m.allgather(arr)
    assert res.shape == (size, 3, 2)
    for p in range(size):
        assert torch.equal(res[p], torch.ones((3, 2), device=device) * p)
    assert torch.equal(_arr, arr)


def test_allgather_jit(device):
    arr = torch.ones((3, 2), device=device) * rank
    f = m.jit(lambda x: m.allgather(x))
    for _ in range(3):
        res = f(arr)
        for p in range(size):
            assert torch.equal(res[p], torch.ones((3, 2), device=device) * p)


def test_allgather_scalar(device):
    res = m.allgather(rank)
    assert res.shape == (size,)
    assert torch.equal(res.cpu(), torch.arange(size))


def test_allgather_scalar_jit(device):
    x = torch.tensor(float(rank), device=device)
    f = m.jit(lambda v: m.allgather(v))
    for _ in range(3):
        assert torch.equal(f(x).cpu(), torch.arange(size, dtype=torch.float32))


def test_allgather_grad(device):
    # extension: adjoint of allgather = sum over ranks of the matching slice
    x = (torch.ones(4, device=device) * (rank + 1)).requires_grad_(True)
    out = m.allgather(x)
    (out * (rank + 1)).sum().backward()
    assert torch.equal(x.grad, torch.ones(4, device=device) * sum(range(1, size + 1)))


def test_allgather_odd_dtypes(device):
    for dt in (torch.uint8, torch.int16, torch.float64, torch.complex64, torch.bool):
        x = torch.ones(5, device=device).to(dt)
        res = m.allgather(x)
        assert res.dtype == dt and res.shape == (size, 5)
        assert torch.equal(res[rank], x)
