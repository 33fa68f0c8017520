"""The CUDA-graph 'jit': caching, pytrees, static args, autograd fall-through."""

import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

size = MPI.COMM_WORLD.Get_size()


def test_jit_pytree_and_static_args(device):
    calls = []

    @m.jit
    def f(state, scale, flag=True):
        calls.append(1)
        a, b = state["a"], state["b"]
        out = m.allreduce(a + b, MPI.SUM) * scale
        return {"sum": out, "pair": (a, out if flag else b)}

    a = torch.ones(5, device=device)
    b = torch.ones(5, device=device) * 2
    for i in range(4):
        res = f({"a": a + i, "b": b}, 2.0)
        assert torch.equal(res["sum"], (a + i + b) * size * 2.0)
        assert torch.equal(res["pair"][0], a + i)
    for _ in range(2):                                # new static arg -> warm-up, then a new graph
        res = f({"a": a, "b": b}, 2.0, flag=False)
        assert torch.equal(res["pair"][1], b)
    if device.type == "cuda":
        assert len(calls) < 8          # replays do not re-run Python
        assert len(f._cache) == 2


def test_jit_falls_through_for_autograd(device):
    f = m.jit(lambda x: m.allreduce(x, MPI.SUM).sum())
    x = torch.ones(3, device=device, requires_grad=True)
    for _ in range(3):
        x.grad = None
        f(x).backward()
        assert torch.equal(x.grad, torch.ones(3, device=device))


def test_jit_donate_outputs(device):
    f = m.jit(lambda x: m.allgather(x), donate_outputs=True)
    x = torch.arange(4, dtype=torch.float32, device=device)
    for _ in range(3):
        assert torch.equal(f(x)[MPI.COMM_WORLD.Get_rank()], x)
