"""Scenario parity with /root/reference/tests/collective_ops/test_allreduce.py (same cases; jax
transforms replaced by their torch / mpi4jax_b200 equivalents)."""

import pytest
import torch
from torch.func import grad, jvp, vjp, vmap

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

comm = MPI.COMM_WORLD
rank = comm.Get_rank()
size = comm.Get_size()


def test_allreduce(device):
    arr = torch.ones((3, 2), device=device)
    _arr = arr.clone()
    res = m.allreduce(arr, op=MPI.SUM)
    assert torch.equal(res, arr * size)
    assert torch.equal(_arr, arr)


def test_allreduce_jit(device):
    arr = torch.ones((3, 2), device=device)
    _arr = arr.clone()
    f = m.jit(lambda x: m.allreduce(x, op=MPI.SUM))
    for _ in range(3):  # eager warm-up, capture, replay
        res = f(arr)
        assert torch.equal(res, arr * size)
    assert torch.equal(_arr, arr)


def test_allreduce_scalar(device):
    res = m.allreduce(1, op=MPI.SUM)
    assert res.shape == ()
    assert res.item() == size


def test_allreduce_scalar_jit(device):
    f = m.jit(lambda x: m.allreduce(x, op=MPI.SUM))
    x = torch.tensor(1.0, device=device)
    for _ in range(3):
        assert f(x).item() == size


def test_allreduce_vmap(device):
    arr = torch.ones((3, 2), device=device)
    _arr = arr.clone()
    res = vmap(lambda x: m.allreduce(x, op=MPI.SUM), in_dims=0, out_dims=0)(arr)
    assert torch.equal(res, arr * size)
    assert torch.equal(_arr, arr)


def test_allreduce_vmap_jit(device):
    arr = torch.ones((3, 2), device=device)
    f = m.jit(vmap(lambda x: m.allreduce(x, op=MPI.SUM), in_dims=0, out_dims=0))
    for _ in range(3):
        assert torch.equal(f(arr), arr * size)


def test_allreduce_transpose(device):
    arr = torch.ones((3, 2), device=device)
    _arr = arr.clone()
    (res,) = m.linear_transpose(lambda x: m.allreduce(x, op=MPI.SUM), arr)(_arr)
    assert torch.equal(_arr, res)


def test_allreduce_transpose_jit(device):
    arr = torch.ones((3, 2), device=device)

    def f(x):
        (res,) = m.linear_transpose(lambda y: m.allreduce(y, op=MPI.SUM), arr)(x)
        return res

    fj = m.jit(f)
    for _ in range(3):
        assert torch.equal(arr, fj(arr))


def test_allreduce_transpose2(device):
    # transposing twice gives the allreduce back
    arr = torch.ones((3, 2), device=device)
    _arr2 = arr.clone()

    def lt(y):
        return m.linear_transpose(lambda x: m.allreduce(x, op=MPI.SUM), arr)(y)[0]

    (res,) = m.linear_transpose(lt, arr)(_arr2)
    expected = m.allreduce(_arr2, op=MPI.SUM)
    assert torch.equal(expected, res)


def test_allreduce_grad(device):
    arr = torch.ones((3, 2), device=device)
    _arr = arr.clone()

    def loss(x):
        return m.allreduce(x, op=MPI.SUM).sum()

    g = grad(loss)(arr)
    assert torch.equal(loss(arr), arr.sum() * size)
    assert torch.equal(_arr, g)

    # classic autograd
    x = arr.clone().requires_grad_(True)
    loss(x).backward()
    assert torch.equal(x.grad, _arr)

    def testfun(x):
        y = m.allreduce(x, op=MPI.SUM)
        z = x + 2 * y  # noqa: F841
        return m.allreduce(x, op=MPI.SUM).sum()

    assert torch.equal(grad(testfun)(arr), _arr)


def test_allreduce_jvp(device):
    arr = torch.ones((3, 2), device=device)
    _arr = arr.clone()
    res, tan = jvp(lambda x: m.allreduce(x, op=MPI.SUM), (arr,), (_arr,))
    assert torch.equal(m.allreduce(arr, op=MPI.SUM), res)
    assert torch.equal(m.allreduce(_arr, op=MPI.SUM), tan)


def test_allreduce_vjp(device):
    arr = torch.ones((3, 2), device=device)
    _arr = arr.clone()
    res, vjp_fun = vjp(lambda x: m.allreduce(x, op=MPI.SUM), arr)
    (ct,) = vjp_fun(_arr)
    assert torch.equal(m.allreduce(arr, op=MPI.SUM), res)
    assert torch.equal(_arr, ct)


def test_allreduce_chained(device):
    def foo(x):
        x1 = m.allreduce(x, op=MPI.SUM, comm=comm)
        x2 = m.allreduce(x, op=MPI.SUM, comm=comm)
        return x1 + x2

    res_t = grad(foo)(torch.tensor(0.0, device=device))
    assert res_t.item() == 2.0


def test_allreduce_non_sum_grad_raises(device):
    x = torch.ones(3, device=device, requires_grad=True)
    y = m.allreduce(x, op=MPI.MAX)
    with pytest.raises(NotImplementedError):
        y.sum().backward()


def test_custom_vjp(device):
    class F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, y):
            ctx.save_for_backward(torch.cos(x), torch.sin(x), y)
            r = (torch.sin(x) * y).sum()
            return m.allreduce(r, op=MPI.SUM)

        @staticmethod
        def backward(ctx, g):
            g = m.allreduce(g, op=MPI.SUM)
            cos_x, sin_x, y = ctx.saved_tensors
            return cos_x * g * y, sin_x * g

    x = torch.ones(3, device=device, requires_grad=True)
    y = (torch.ones(3, device=device) * 2).requires_grad_(True)
    out = F.apply(x, y)
    out.backward()
    assert torch.allclose(x.grad, torch.cos(torch.ones(3, device=device)) * 2 * size)


def test_advanced_expectation_gradient(device):
    """Analogue of the reference's netket-style test (test_allreduce.py:252-322): a custom
    backward that itself differentiates through allreduce + vmap."""
    torch.manual_seed(3)
    w = torch.randn(4, 8, device=device, dtype=torch.float64)
    x = torch.randn(16, 4, device=device, dtype=torch.float64)

    def log_pdf(w, x):
        return (x @ w).sum(-1)

    def expected_fun(w, x):
        return torch.exp((x @ w).sum(-1)) - 2

    class Expect(torch.autograd.Function):
        @staticmethod
        def forward(ctx, w):
            L = expected_fun(w, x)
            mean = m.allreduce(L.mean(), op=MPI.SUM) / size
            ctx.save_for_backward(w, L - mean)
            return mean

        @staticmethod
        def backward(ctx, dout):
            w0, dL = ctx.saved_tensors

            def f(w_):
                term1 = vmap(torch.mul)(dL, log_pdf(w_, x))
                term2 = expected_fun(w_, x)
                return (m.allreduce((term1 + term2).mean(0), op=MPI.SUM) / size).sum()

            _, pb = vjp(f, w0)
            return pb(dout)[0]

    wv = w.clone().requires_grad_(True)
    out = Expect.apply(wv)
    out.backward()
    assert wv.grad.shape == w.shape
    assert torch.isfinite(wv.grad).all()


@pytest.mark.parametrize("opname", ["SUM", "PROD", "MIN", "MAX"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.bfloat16, torch.float16,
                                   torch.int32, torch.int64, torch.uint8])
def test_allreduce_ops_dtypes(device, opname, dtype):
    """Gap the reference leaves open: ops other than SUM, more dtypes."""
    op = getattr(MPI, opname)
    base = torch.arange(1, 7).reshape(3, 2)
    arr = (base + rank).to(dtype).to(device)
    res = m.allreduce(arr, op=op)
    stack = torch.stack([(base + r).to(torch.float64) for r in range(size)])
    if opname == "SUM":
        exp = stack.sum(0)
    elif opname == "PROD":
        exp = stack.prod(0)
    elif opname == "MIN":
        exp = stack.amin(0)
    else:
        exp = stack.amax(0)
    assert res.dtype == dtype
    assert torch.allclose(res.to(torch.float64).cpu(), exp.to(dtype).to(torch.float64), rtol=1e-2)


@pytest.mark.parametrize("opname", ["LAND", "LOR", "LXOR", "BAND", "BOR", "BXOR"])
def test_allreduce_logical_bitwise(device, opname):
    op = getattr(MPI, opname)
    arr = (torch.tensor([0, 1, 2, 3, 0xF0, 0x0F], dtype=torch.int32) + rank).to(device)
    res = m.allreduce(arr, op=op).cpu()
    vals = [torch.tensor([0, 1, 2, 3, 0xF0, 0x0F], dtype=torch.int32) + r for r in range(size)]
    exp = vals[0].clone()
    for v in vals[1:]:
        if opname == "LAND":
            exp = ((exp != 0) & (v != 0)).to(torch.int32)
        elif opname == "LOR":
            exp = ((exp != 0) | (v != 0)).to(torch.int32)
        elif opname == "LXOR":
            exp = ((exp != 0) ^ (v != 0)).to(torch.int32)
        elif opname == "BAND":
            exp = exp & v
        elif opname == "BOR":
            exp = exp | v
        else:
            exp = exp ^ v
    if size == 1 and opname in ("LAND", "LOR", "LXOR"):
        exp = arr.cpu()       # a single contribution passes through unchanged
    assert torch.equal(res, exp)


def test_allreduce_bool_and_complex(device):
    b = torch.tensor([True, False, rank % 2 == 0], device=device)
    assert m.allreduce(b, op=MPI.LOR).dtype == torch.bool
    c = torch.full((4,), 1 + 2j, dtype=torch.complex64, device=device)
    assert torch.allclose(m.allreduce(c, op=MPI.SUM), c * size)


def test_allreduce_large_and_odd_sizes(device):
    for n in (1, 3, 17, 1000, 4099, 70001, 300_000):
        x = torch.arange(n, dtype=torch.float32, device=device) + rank
        exp = torch.arange(n, dtype=torch.float32, device=device) * size + sum(range(size))
        assert torch.equal(m.allreduce(x, op=MPI.SUM), exp), n


def test_token_rejected(device):
    with pytest.raises(RuntimeError, match="Explicit token management is not supported"):
        m.allreduce(torch.ones(2, device=device), op=MPI.SUM, token=object())


def test_op_type_checked(device):
    with pytest.raises(TypeError, match='unexpected type for argument "op"'):
        m.allreduce(torch.ones(2, device=device), op="sum")
