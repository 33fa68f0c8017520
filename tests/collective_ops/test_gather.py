"""Ports of /root/reference/tests/collective_ops/test_gather.py."""

import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

comm = MPI.COMM_WORLD
rank = comm.Get_rank()
size = comm.Get_size()


def _check(res, arr, device):
    if rank == 0:
        assert res.shape == (size, *arr.shape)
        for p in range(size):
            assert torch.equal(res[p], torch.ones_like(arr) * p)
    else:
        assert torch.equal(res, arr)


def test_gather(device):
    arr = torch.ones((3, 2), device=device) * rank
    _check(m.gather(arr, root=0), arr, device)


def test_gather_jit(device):
    arr = torch.ones((3, 2), device=device) * rank
    f = m.jit(lambda x: m.gather(x, root=0))
    for _ in range(3):
        _check(f(arr), arr, device)


def test_gather_scalar(device):
    res = m.gather(rank, root=0)
    if rank == 0:
        assert torch.equal(res.cpu(), torch.arange(size))
    else:
        assert res.item() == rank


def test_gather_nonzero_root(device):
    root = size - 1
    arr = torch.ones(4, device=device) * rank
    res = m.gather(arr, root=root)
    if rank == root:
        assert torch.equal(res[:, 0].cpu(), torch.arange(size, dtype=torch.float32))
    else:
        assert torch.equal(res, arr)
