"""bcast: everybody ends up with the root's array, the root gets its own input back (scenario parity
with /root/reference/tests/collective_ops/test_bcast.py; plus other roots and the VJP)."""

import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

comm = MPI.COMM_WORLD
rank = comm.Get_rank()
size = comm.Get_size()


def test_bcast(device):
    arr = torch.ones((3, 2), device=device)
    _arr = arr.clone()
    if rank != 0:
        _arr = _arr * 0
    res = m.bcast(_arr, root=0)
    assert torch.equal(res, arr)
    if rank == 0:
        assert torch.equal(_arr, arr)


def test_bcast_jit(device):
    arr = torch.ones((3, 2), device=device)
    _arr = arr.clone()
    if rank != 0:
        _arr = _arr * 0
    f = m.jit(lambda x: m.bcast(x, root=0))
    for _ in range(3):
        assert torch.equal(f(_arr), arr)


def test_bcast_scalar(device):
    _arr = 1 if rank == 0 else 0
    res = m.bcast(_arr, root=0)
    assert res.item() == 1


def test_bcast_nonzero_root(device):
    root = size - 1
    x = torch.full((5,), float(rank), device=device)
    res = m.bcast(x, root=root)
    assert torch.equal(res, torch.full((5,), float(root), device=device))


def test_bcast_grad(device):
    """Extension: VJP of bcast = reduce(SUM) to the root."""
    x = torch.ones(3, device=device, requires_grad=True)
    y = m.bcast(x, root=0)
    (y * (rank + 1)).sum().backward()
    if rank == 0:
        assert torch.equal(x.grad, torch.ones(3, device=device) * sum(range(1, size + 1)))
    else:
        assert torch.equal(x.grad, torch.zeros(3, device=device))
