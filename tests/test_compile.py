"""torch.library registration: the ops are traceable custom ops with an ORDERED effect, the
torch counterpart of the reference's JAX primitives + ordered effect token
(/root/reference/mpi4jax/_src/utils.py:45-53; acceptance test = test_send_recv_deadlock)."""

import os

import pytest
import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI
from mpi4jax_b200 import compiled as mc
from mpi4jax_b200._src import compile_ops

comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()


def test_ops_are_registered_and_effectful():
    for name in compile_ops.ALL_OPS:
        assert hasattr(torch.ops.mpi4jax_b200, name)
    assert compile_ops.ORDERED_EFFECT, "ordered effect registration failed (torch internals moved?)"
    from torch._higher_order_ops.effects import _EffectType, _get_effect

    assert _get_effect("mpi4jax_b200::send") == _EffectType.ORDERED
    assert _get_effect("mpi4jax_b200::barrier") == _EffectType.ORDERED


def test_fake_kernels_give_reference_shapes():
    """abstract-eval rules of the reference: allgather S -> (nproc, *S), ..."""
    from torch._subclasses.fake_tensor import FakeTensorMode

    with FakeTensorMode():
        x = torch.empty(3, 2)
        assert mc.allreduce(x, MPI.SUM, comm=comm).shape == (3, 2)
        assert mc.allgather(x, comm=comm).shape == (size, 3, 2)
        assert mc.alltoall(torch.empty(size, 4), comm=comm).shape == (size, 4)
        assert mc.recv(x, source=0, comm=comm).shape == (3, 2)
        assert mc.sendrecv(torch.empty(5), torch.empty(7), 0, 0, comm=comm).shape == (7,)
        assert mc.send(x, 0, comm=comm) is None


def test_compiled_allreduce_and_grad(device):
    f = torch.compile(lambda x: (mc.allreduce(x * 2, MPI.SUM, comm=comm) + 1).sum(),
                      backend="aot_eager", fullgraph=True)
    x = torch.ones(4, device=device, requires_grad=True)
    y = f(x)
    assert y.item() == (2 * size + 1) * 4
    y.backward()
    assert torch.equal(x.grad, torch.full((4,), 2.0, device=device))


def test_compiled_program_order_is_kept(device):
    """send / recv / barrier have no data dependence on each other: only the ordered effect
    keeps them in program order (and keeps the result-less send alive)."""

    def exchange(arr):
        other = (rank + 1) % size
        if size == 1:
            mc.send(arr, rank, tag=3, comm=comm)
            mc.barrier(comm=comm)
            return mc.recv(arr, source=rank, tag=3, comm=comm)
        if rank % 2 == 0:
            mc.send(arr, other, comm=comm)
            return mc.recv(arr, (rank - 1) % size, comm=comm)
        got = mc.recv(arr, (rank - 1) % size, comm=comm)
        mc.send(arr, other, comm=comm)
        return got

    if size > 1 and size % 2:
        pytest.skip("ring with an odd number of ranks needs sendrecv")
    f = torch.compile(exchange, backend="aot_eager", fullgraph=True)
    arr = torch.ones(10, device=device) * rank
    for _ in range(2):
        out = f(arr)
        assert torch.equal(out, torch.ones(10, device=device) * ((rank - 1) % size))
    m.flush()


def test_torch_export_keeps_all_ops(device):
    """torch.export produces a graph that contains every op, including the result-less barrier."""

    class Mod(torch.nn.Module):
        def forward(self, x):
            y = mc.allreduce(x * 2, MPI.SUM, comm=comm)
            mc.barrier(comm=comm)
            return mc.allgather(y, comm=comm)

    ep = torch.export.export(Mod(), (torch.ones(3, device=device),))
    targets = [str(n.target) for n in ep.graph.nodes if n.op == "call_function"]
    for name in ("allreduce", "barrier", "allgather"):
        assert any(f"mpi4jax_b200.{name}" in t for t in targets), targets
    out = ep.module()(torch.ones(3, device=device))
    assert torch.equal(out, torch.full((size, 3), 2.0 * size, device=device))


def test_compiled_gradients_match_eager_rules(device):
    """allgather / alltoall / bcast / sendrecv are differentiable in the traced frontend too, with
    the same adjoints as the eager ops."""
    x = torch.arange(size * 3, dtype=torch.float32, device=device).reshape(size, 3) + rank

    def via(ns, t):
        a = ns.allgather(t[0], comm=comm)                       # (size, 3)
        b = ns.alltoall(t * 2, comm=comm)                       # (size, 3)
        c = ns.bcast(t[0] * 3, 0, comm=comm)                    # (3,)
        d = ns.sendrecv(t[0] * 5, t[0], (rank - 1) % size, (rank + 1) % size, comm=comm)
        w = torch.arange(1, 4, dtype=torch.float32, device=t.device)
        return (a * w).sum() + (b * b).sum() + (c * w).sum() + (d * w).sum()

    xe = x.clone().requires_grad_(True)
    via(m, xe).backward()
    xc = x.clone().requires_grad_(True)
    f = torch.compile(lambda t: via(mc, t), backend="aot_eager", fullgraph=True)
    f(xc).backward()
    assert torch.allclose(xc.grad, xe.grad)
    m.flush()


def test_compiled_gather_scatter_allgather_shapes_under_compile(device):
    """Rank-dependent static arguments (root / is_root, communicator size) are resolved while
    tracing; no registry lookups inside the traced frame."""

    def f(t):
        g = mc.gather(t, 0, comm=comm)
        a = mc.allgather(t, comm=comm)
        s = mc.scatter(a if rank == 0 else t, 0, comm=comm)
        return g, a, s

    t = torch.arange(4, dtype=torch.float32, device=device) + 10 * rank
    g, a, s = torch.compile(f, backend="aot_eager", fullgraph=True)(t)
    assert a.shape == (size, 4) and torch.equal(a[rank], t)
    assert (g.shape == (size, 4)) if rank == 0 else torch.equal(g, t)
    assert torch.equal(s, torch.arange(4, dtype=torch.float32, device=device) + 10 * rank)
    m.flush()


def test_compiled_recv_and_sendrecv_fill_a_status(device):
    """``status=`` on the traceable path (reference: the Status is filled under jit,
    tests/collective_ops/test_send_and_recv.py:113-153, test_sendrecv.py:28-60)."""
    st_a, st_b = MPI.Status(), MPI.Status()
    nxt, prv = (rank + 1) % size, (rank - 1) % size

    def f(t):
        a = mc.sendrecv(t, t, prv, nxt, sendtag=4, recvtag=4, comm=comm, status=st_a)
        mc.send(t * 2, nxt, tag=9, comm=comm)
        b = mc.recv(t, prv, tag=9, comm=comm, status=st_b)
        return a, b

    t = torch.arange(6, dtype=torch.float32, device=device) + 100 * rank
    a, b = torch.compile(f, backend="aot_eager", fullgraph=True)(t)
    want = torch.arange(6, dtype=torch.float32, device=device) + 100 * prv
    assert torch.equal(a, want) and torch.equal(b, want * 2)
    for st, tag in ((st_a, 4), (st_b, 9)):
        assert st.Get_source() == prv and st.Get_tag() == tag and st.Get_count() == 6
    m.flush()
