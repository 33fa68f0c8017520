"""scan: inclusive prefix reduction over the ranks 0..r, same shape and dtype as the input
(reference: /root/reference/mpi4jax/_src/collective_ops/scan.py:44-60, 113-114)."""

import math

import pytest
import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI


comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()


def _prefix(op, values):
    acc = values[0]
    for v in values[1:]:
        acc = op(acc, v)
    return acc


CASES = [
    # op, per-rank scalar as a function of the rank, python combiner, dtype
    ("SUM", lambda r: r, lambda a, b: a + b, torch.float32),
    ("PROD", lambda r: r + 1, lambda a, b: a * b, torch.float32),
    ("MAX", lambda r: (7 * r) % 5, max, torch.int64),
    ("MIN", lambda r: 10 - r, min, torch.float64),
    ("BOR", lambda r: 1 << (r % 8), lambda a, b: a | b, torch.int32),
]


@pytest.mark.parametrize("name, gen, combine, dtype", CASES, ids=[c[0] for c in CASES])
def test_prefix_over_ranks(device, name, gen, combine, dtype):
    x = torch.full((3, 2), gen(rank), dtype=dtype, device=device)
    keep = x.clone()
    out = m.scan(x, op=getattr(MPI, name))
    want = _prefix(combine, [gen(r) for r in range(rank + 1)])
    assert out.dtype == dtype and out.shape == x.shape
    assert torch.equal(out, torch.full((3, 2), want, dtype=dtype, device=device))
    assert torch.equal(x, keep)


def test_elementwise_not_across_elements(device):
    x = torch.arange(6, dtype=torch.float32, device=device) * (rank + 1)
    out = m.scan(x, op=MPI.SUM)
    tri = (rank + 1) * (rank + 2) // 2                       # 1 + 2 + ... + (rank + 1)
    assert torch.equal(out, torch.arange(6, dtype=torch.float32, device=device) * tri)


def test_python_int_is_accepted(device):
    assert m.scan(rank, op=MPI.SUM).item() == rank * (rank + 1) // 2


def test_python_float_is_accepted(device):
    assert math.isclose(m.scan(0.5, op=MPI.SUM).item(), 0.5 * (rank + 1))


def test_replay_under_jit(device):
    x = torch.ones((3, 2), device=device) * rank
    running = m.jit(lambda t: m.scan(t, op=MPI.SUM))
    for _ in range(3):
        assert torch.equal(running(x), torch.ones((3, 2), device=device) * (rank * (rank + 1) // 2))
