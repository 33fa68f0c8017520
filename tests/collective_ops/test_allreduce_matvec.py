"""Linear-algebra property suite: column-sharded mat-vec with allreduce, nested transposes,
jvp/vjp -- scenario parity with /root/reference/tests/collective_ops/test_allreduce_matvec.py."""

import pytest
import torch
from torch.func import jvp, vjp

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

comm = MPI.COMM_WORLD
rank = comm.Get_rank()
size = comm.Get_size()

M, N = 5, 60
pytestmark = pytest.mark.skipif(N % size != 0, reason="60 % size must be 0")


class Problem:
    def __init__(self, device):
        gen = torch.Generator().manual_seed(123)
        self.A = torch.rand(M, N, generator=gen, dtype=torch.float64).to(device)
        self.x = torch.rand(N, generator=gen, dtype=torch.float64).to(device)
        self.y = torch.rand(M, generator=gen, dtype=torch.float64).to(device)
        self.v = torch.rand(N, generator=gen, dtype=torch.float64).to(device)
        self.vprime = torch.rand(M, generator=gen, dtype=torch.float64).to(device)
        n_local = N // size
        lo, hi = n_local * rank, n_local * (rank + 1)
        self.A_local = self.A[:, lo:hi].contiguous()
        self.x_local = self.x[lo:hi].contiguous()
        self.v_local = self.v[lo:hi].contiguous()
        self.Ax = self.A @ self.x
        self.ATy_local = (self.A.T @ self.y)[lo:hi]
        self.Av = self.A @ self.v
        self.ATvprime_local = (self.A.T @ self.vprime)[lo:hi]

    def mv(self, x_local):
        return m.allreduce(self.A_local @ x_local, op=MPI.SUM, comm=comm)

    def mvT(self, y):
        ar_T = transpose(lambda t: m.allreduce(t, op=MPI.SUM, comm=comm), y)
        return self.A_local.T @ ar_T(y)


def transpose(f, x):
    def fT(y):
        return m.linear_transpose(f, x)(y)[0]

    return fT


@pytest.fixture
def P(device):
    return Problem(device)


def both(fn):
    """eager and CUDA-graph ('jit') variants, like the reference's *_jit twins"""
    return [fn, m.jit(fn)]


def test_matvec(P):
    for f in both(P.mv):
        for _ in range(3):
            assert torch.allclose(P.Ax, f(P.x_local))


def test_matvecT(P):
    for f in both(P.mvT):
        for _ in range(3):
            assert torch.allclose(P.ATy_local, f(P.y))


def test_matvec_transpose(P):
    for f in both(transpose(P.mv, P.x_local)):
        for _ in range(3):
            assert torch.allclose(P.ATy_local, f(P.y))


def test_matvecT_transpose(P):
    for f in both(transpose(P.mvT, P.y)):
        for _ in range(3):
            assert torch.allclose(P.Ax, f(P.x_local))


def test_matvec_transpose_transpose(P):
    ltlt = transpose(transpose(P.mv, P.x_local), P.y)
    for f in both(ltlt):
        for _ in range(3):
            assert torch.allclose(P.Ax, f(P.x_local))


def test_matvecT_transpose_transpose(P):
    ltlt = transpose(transpose(P.mvT, P.y), P.x_local)
    for f in both(ltlt):
        for _ in range(3):
            assert torch.allclose(P.ATy_local, f(P.y))


def test_matvec_transpose3(P):
    lt3 = transpose(transpose(transpose(P.mv, P.x_local), P.y), P.x_local)
    assert torch.allclose(P.ATy_local, lt3(P.y))


def test_matvec_jvp(P):
    res, tan = jvp(P.mv, (P.x_local,), (P.v_local,))
    assert torch.allclose(P.Ax, res)
    assert torch.allclose(P.Av, tan)


def test_matvec_vjp(P):
    res, vjp_fun = vjp(P.mv, P.x_local)
    (ct,) = vjp_fun(P.vprime)
    assert torch.allclose(P.Ax, res)
    assert torch.allclose(P.ATvprime_local, ct)


def test_matvecT_jvp(P):
    res, tan = jvp(P.mvT, (P.y,), (P.vprime,))
    assert torch.allclose(P.ATy_local, res)
    assert torch.allclose(P.ATvprime_local, tan)


def test_matvecT_vjp(P):
    res, vjp_fun = vjp(P.mvT, P.y)
    (ct,) = vjp_fun(P.v_local)
    assert torch.allclose(P.ATy_local, res)
    assert torch.allclose(P.Av, ct)
