"""Ports of /root/reference/tests/collective_ops/test_scatter.py."""

import pytest
import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

comm = MPI.COMM_WORLD
rank = comm.Get_rank()
size = comm.Get_size()


def _input(device):
    if rank == 0:
        return torch.stack([torch.ones((3, 2)) * r for r in range(size)]).to(device)
    return torch.ones((3, 2), device=device) * rank


def test_scatter(device):
    res = m.scatter(_input(device), root=0)
    assert torch.equal(res, torch.ones((3, 2), device=device) * rank)


def test_scatter_jit(device):
    x = _input(device)
    f = m.jit(lambda v: m.scatter(v, root=0))
    for _ in range(3):
        assert torch.equal(f(x), torch.ones((3, 2), device=device) * rank)


def test_scatter_wrong_size(device):
    if rank == 0:
        with pytest.raises(ValueError) as excinfo:
            m.scatter(torch.ones((size + 1, 3, 2), device=device), root=0)
        assert "Scatter input must have shape (nproc, ...)" in str(excinfo.value)
