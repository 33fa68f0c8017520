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


def test_vmap_allgather_alltoall_bcast(device):
    """vmap rules beyond the reference's (allreduce, barrier, sendrecv): the batch travels as one
    message; results equal a Python loop over the batch."""
    from torch.func import vmap

    comm = MPI.COMM_WORLD
    rank, size = comm.Get_rank(), comm.Get_size()
    xb = torch.arange(3 * 4, dtype=torch.float32, device=device).reshape(3, 4) + 100 * rank

    got = vmap(lambda x: m.allgather(x, comm=comm))(xb)
    want = torch.stack([m.allgather(x, comm=comm) for x in xb])
    assert got.shape == (3, size, 4) and torch.equal(got, want)

    got = vmap(lambda x: m.allgather(x, comm=comm), in_dims=1, out_dims=1)(xb)       # batch over columns
    want = torch.stack([m.allgather(xb[:, j], comm=comm) for j in range(4)], dim=1)
    assert got.shape == (size, 4, 3) and torch.equal(got, want)

    ab = torch.arange(5 * size * 2, dtype=torch.float32, device=device).reshape(5, size, 2) + 1000 * rank
    got = vmap(lambda x: m.alltoall(x, comm=comm))(ab)
    want = torch.stack([m.alltoall(x, comm=comm) for x in ab])
    assert got.shape == (5, size, 2) and torch.equal(got, want)

    got = vmap(lambda x: m.bcast(x, 0, comm=comm))(xb)
    want = torch.stack([m.bcast(x, 0, comm=comm) for x in xb])
    assert torch.equal(got, want)
    if rank != 0:
        assert torch.equal(got, torch.arange(12, dtype=torch.float32, device=device).reshape(3, 4))
    # and through grad-of-vmap: d/dx sum(allgather(x)) = nproc
    g = torch.func.grad(lambda x: vmap(lambda r: m.allgather(r, comm=comm))(x).sum())(xb)
    assert torch.equal(g, torch.full_like(xb, float(size)))


def test_func_grad_through_bcast(device):
    """bcast VJP (= reduce of the cotangents to the root) under torch.func.grad and vmap-of-grad."""
    comm = MPI.COMM_WORLD
    rank, size = comm.Get_rank(), comm.Get_size()
    x = torch.arange(4, dtype=torch.float32, device=device) + 1

    def loss(t):
        return (m.bcast(t * 2, 0, comm=comm) ** 2).sum()

    g = torch.func.grad(loss)(x)
    want = 8 * x * size if rank == 0 else torch.zeros_like(x)
    assert torch.allclose(g, want)
    gb = torch.func.vmap(torch.func.grad(loss))(torch.stack([x, 2 * x]))
    assert torch.allclose(gb, torch.stack([want, 2 * want]))
