"""Ports of /root/reference/tests/collective_ops/test_reduce.py."""

import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

comm = MPI.COMM_WORLD
rank = comm.Get_rank()
size = comm.Get_size()


def test_reduce(device):
    arr = torch.ones((3, 2), device=device) * rank
    _arr = arr.clone()
    res = m.reduce(arr, op=MPI.SUM, root=0)
    if rank == 0:
        assert torch.equal(res, torch.ones((3, 2), device=device) * sum(range(size)))
    else:
        assert torch.equal(res, arr)
    assert torch.equal(_arr, arr)


def test_reduce_jit(device):
    arr = torch.ones((3, 2), device=device) * rank
    f = m.jit(lambda x: m.reduce(x, op=MPI.SUM, root=0))
    for _ in range(3):
        res = f(arr)
        if rank == 0:
            assert torch.equal(res, torch.ones((3, 2), device=device) * sum(range(size)))
        else:
            assert torch.equal(res, arr)


def test_reduce_scalar(device):
    res = m.reduce(rank, op=MPI.SUM, root=0)
    assert res.item() == (sum(range(size)) if rank == 0 else rank)


def test_reduce_max_nonzero_root(device):
    root = size - 1
    arr = torch.arange(5, dtype=torch.float32, device=device) + rank
    res = m.reduce(arr, op=MPI.MAX, root=root)
    if rank == root:
        assert torch.equal(res, torch.arange(5, dtype=torch.float32, device=device) + size - 1)
    else:
        assert torch.equal(res, arr)
