"""Ports of /root/reference/tests/collective_ops/test_scan.py."""

import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

comm = MPI.COMM_WORLD
rank = comm.Get_rank()
size = comm.Get_size()


def test_scan(device):
    arr = torch.ones((3, 2), device=device) * rank
    res = m.scan(arr, op=MPI.SUM)
    assert torch.equal(res, torch.ones((3, 2), device=device) * sum(range(rank + 1)))


def test_scan_jit(device):
    arr = torch.ones((3, 2), device=device) * rank
    f = m.jit(lambda x: m.scan(x, op=MPI.SUM))
    for _ in range(3):
        assert torch.equal(f(arr), torch.ones((3, 2), device=device) * sum(range(rank + 1)))


def test_scan_scalar(device):
    assert m.scan(rank, op=MPI.SUM).item() == sum(range(rank + 1))


def test_scan_prod(device):
    arr = torch.full((4,), float(rank + 1), device=device)
    exp = 1.0
    for r in range(rank + 1):
        exp *= r + 1
    assert torch.equal(m.scan(arr, op=MPI.PROD), torch.full((4,), exp, device=device))
