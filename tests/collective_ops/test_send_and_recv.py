"""Ports of /root/reference/tests/collective_ops/test_send_and_recv.py (+ ANY_SOURCE/ANY_TAG,
self-send and large-message cases the reference does not cover)."""

import pytest
import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

comm = MPI.COMM_WORLD
rank = comm.Get_rank()
size = comm.Get_size()

need2 = pytest.mark.skipif(size < 2, reason="need at least 2 processes to test send/recv")


@need2
def test_send_recv(device):
    arr = torch.ones((3, 2), device=device) * rank
    _arr = arr.clone()
    if rank == 0:
        for proc in range(1, size):
            res = m.recv(arr, source=proc, tag=proc)
            assert torch.equal(res, torch.ones_like(arr) * proc)
            assert torch.equal(_arr, arr)
    else:
        m.send(arr, 0, tag=rank)
        assert torch.equal(_arr, arr)


@need2
def test_send_recv_scalar(device):
    if rank == 0:
        for proc in range(1, size):
            res = m.recv(0, source=proc, tag=proc)
            assert res.item() == proc
    else:
        m.send(rank, 0, tag=rank)


@need2
def test_send_recv_jit(device):
    arr = torch.ones((3, 2), device=device) * rank

    @m.jit
    def send_jit(x):
        m.send(x, 0, tag=rank)
        return x

    for _ in range(3):
        if rank == 0:
            for proc in range(1, size):
                res = m.jit(lambda x: m.recv(x, source=proc, tag=proc))(arr)
                assert torch.equal(res, torch.ones_like(arr) * proc)
        else:
            send_jit(arr)


@pytest.mark.skipif(size < 2 or rank > 1, reason="Runs only on rank 0 and 1")
def test_send_recv_deadlock(device):
    """Ops must execute in program order (reference: hangs if XLA reorders the custom calls)."""

    @m.jit
    def exchange(arr):
        if rank == 0:
            m.send(arr, 1)
            newarr = m.recv(arr, 1)
        else:
            newarr = m.recv(arr, 0)
            m.send(arr, 0)
        return newarr

    arr = torch.ones(10, device=device) * rank
    for _ in range(3):
        out = exchange(arr)
        assert torch.equal(out, torch.ones_like(arr) * (1 - rank))


@need2
def test_send_recv_status(device):
    arr = torch.ones((3, 2), device=device) * rank
    if rank == 0:
        for proc in range(1, size):
            status = MPI.Status()
            res = m.recv(arr, source=proc, tag=proc, status=status)
            assert torch.equal(res, torch.ones_like(arr) * proc)
            assert status.Get_source() == proc
            assert status.Get_tag() == proc
            assert status.Get_count() == 6
    else:
        m.send(arr, 0, tag=rank)


@need2
def test_recv_any_source_any_tag(device):
    arr = torch.ones(4, device=device) * rank
    if rank == 0:
        seen = set()
        for _ in range(1, size):
            status = MPI.Status()
            res = m.recv(arr, status=status)           # ANY_SOURCE, ANY_TAG
            src = status.Get_source()
            assert torch.equal(res, torch.ones_like(arr) * src)
            assert status.Get_tag() == 100 + src
            seen.add(src)
        assert seen == set(range(1, size))
    else:
        m.send(arr, 0, tag=100 + rank)


def test_send_recv_self(device):
    arr = torch.arange(7, dtype=torch.float32, device=device)
    m.send(arr, rank, tag=5)
    res = m.recv(torch.empty_like(arr), source=rank, tag=5)
    assert torch.equal(res, arr)


@need2
def test_send_recv_large(device):
    n = 3_000_001        # several ring fragments on the GPU path, odd size
    other = (rank + 1) % size
    prev = (rank - 1) % size
    arr = torch.arange(n, dtype=torch.float32, device=device) + rank
    if rank % 2 == 0:
        m.send(arr, other)
        res = m.recv(arr, prev)
    else:
        res = m.recv(arr, prev)
        m.send(arr, other)
    if size % 2 == 0:
        assert torch.equal(res, torch.arange(n, dtype=torch.float32, device=device) + prev)


@need2
def test_send_recv_grad(device):
    """Extension: recv is differentiable (adjoint = send back), send_with_grad closes the loop."""
    if rank == 0:
        x = torch.ones(3, device=device, requires_grad=True)
        tok = m.send_with_grad(x * 2, 1)
        tok.backward()
        assert torch.equal(x.grad, torch.ones(3, device=device) * 2 * 3)
    elif rank == 1:
        y = m.recv(torch.empty(3, device=device, requires_grad=True), source=0)
        (y * 3).sum().backward()
