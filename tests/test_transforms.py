"""Integration with host-framework control flow: conjugate gradients whose operator is an
allreduce (port of /root/reference/tests/test_jax_transforms.py, where
jax.scipy.sparse.linalg.cg runs a while_loop with mpi4jax effects inside)."""

import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

comm = MPI.COMM_WORLD


def cg(matvec, b, iters=25):
    x = torch.zeros_like(b)
    r = b - matvec(x)
    p = r.clone()
    rs = (r * r).sum()
    for _ in range(iters):
        Ap = matvec(p)
        alpha = rs / (p * Ap).sum().clamp_min(1e-30)
        x = x + alpha * p
        r = r - alpha * Ap
        rs_new = (r * r).sum()
        p = r + (rs_new / rs.clamp_min(1e-30)) * p
        rs = rs_new
    return x


def test_custom_linear_solver(device):
    gen = torch.Generator().manual_seed(1)
    b = torch.randn(24, generator=gen, dtype=torch.float64).to(device)

    def mat_vec(v):
        return m.allreduce(v, op=MPI.SUM, comm=comm)

    x = cg(mat_vec, b)
    assert torch.allclose(comm.Get_size() * x, b)
    solver = m.jit(lambda rhs: cg(mat_vec, rhs))     # the whole solve as one CUDA graph
    for _ in range(3):
        x = solver(b)
        assert torch.allclose(comm.Get_size() * x, b)


def test_ops_in_python_loop_jit(device):
    """fori_loop analogue: ops inside a loop inside one jitted function."""

    def f(x):
        for _ in range(5):
            x = m.allreduce(x, op=MPI.SUM) / comm.Get_size() + 1
            m.barrier()
        return x

    fj = m.jit(f)
    x = torch.zeros(8, device=device)
    for _ in range(3):
        assert torch.equal(fj(x), torch.full((8,), 5.0, device=device))
