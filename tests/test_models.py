"""MLP through the differentiable ops (BASELINE config 5: grad through allreduce + bcast)."""

import pytest
import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI
from mpi4jax_b200.models import ParallelMLP

comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()


def _dense_reference(xs, ys, d_in, d_hidden, d_out, lr):
    """Single-process reference: the same weights, the concatenated batch."""
    gen = torch.Generator().manual_seed(0)
    w1 = (torch.randn(d_in, d_hidden, generator=gen) / d_in**0.5).double().requires_grad_(True)
    w2 = (torch.randn(d_hidden, d_out, generator=gen) / d_hidden**0.5).double().requires_grad_(True)
    loss = sum(((torch.tanh(x @ w1) @ w2 - y) ** 2).mean() for x, y in zip(xs, ys)) / len(xs)
    loss.backward()
    return loss.detach(), w1 - lr * w1.grad, w2 - lr * w2.grad


def test_mlp_data_parallel_grad_matches_dense(device):
    d_in, d_h, d_out, lr = 6, 8, 3, 0.1
    gens = [torch.Generator().manual_seed(100 + r) for r in range(size)]
    xs = [torch.randn(5, d_in, generator=g, dtype=torch.float64) for g in gens]
    ys = [torch.randn(5, d_out, generator=g, dtype=torch.float64) for g in gens]
    mlp = ParallelMLP(d_in, d_h, d_out, comm=comm, device=device, dtype=torch.float64, mode="dp")
    loss = mlp.step(xs[rank].to(device), ys[rank].to(device), lr=lr)
    ref_loss, w1, w2 = _dense_reference(xs, ys, d_in, d_h, d_out, lr)
    assert torch.allclose(loss.cpu(), ref_loss)
    if rank == 0:     # the bcast VJP delivers the summed gradient to the root
        assert torch.allclose(mlp.w1.detach().cpu(), w1.detach())
        assert torch.allclose(mlp.w2.detach().cpu(), w2.detach())


@pytest.mark.skipif(8 % size != 0, reason="hidden size 8 must divide")
def test_mlp_tensor_parallel_matches_dense(device):
    d_in, d_h, d_out, lr = 6, 8, 3, 0.1
    g = torch.Generator().manual_seed(7)
    x = torch.randn(5, d_in, generator=g, dtype=torch.float64)
    y = torch.randn(5, d_out, generator=g, dtype=torch.float64)
    mlp = ParallelMLP(d_in, d_h, d_out, comm=comm, device=device, dtype=torch.float64, mode="tp")
    loss = mlp.step(x.to(device), y.to(device), lr=lr)
    ref_loss, w1, w2 = _dense_reference([x], [y], d_in, d_h, d_out, lr)
    k = d_h // size
    assert torch.allclose(loss.cpu(), ref_loss)
    assert torch.allclose(mlp.w1.detach().cpu(), w1.detach()[:, rank * k:(rank + 1) * k])
    assert torch.allclose(mlp.w2.detach().cpu(), w2.detach()[rank * k:(rank + 1) * k])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["dp", "tp"])
def test_mlp_training_step_as_one_cuda_graph(mode):
    """mpi4jax_b200.jit(mlp.step): forward, backward through the collectives' VJPs and the update are
    captured into ONE CUDA graph (the reference's headline use: jax.grad of a jitted loss that
    calls allreduce / bcast, README.rst:59-96); replays match eager steps of a twin model."""
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    if mode == "tp" and 8 % size != 0:
        pytest.skip("hidden size 8 must divide")
    dev = comm.device
    g = torch.Generator().manual_seed(11 + (rank if mode == "dp" else 0))
    x = torch.randn(5, 6, generator=g).to(dev)
    y = torch.randn(5, 3, generator=g).to(dev)
    a = ParallelMLP(6, 8, 3, comm=comm, device=dev, dtype=torch.float32, mode=mode)
    b = ParallelMLP(6, 8, 3, comm=comm, device=dev, dtype=torch.float32, mode=mode)
    with torch.no_grad():
        for pa, pb in zip(a.parameters(), b.parameters()):
            pb.copy_(pa)
    fast = m.jit(lambda u, v: b.step(u, v, lr=0.1))
    for _ in range(4):                      # call 1 warms up eagerly, call 2 captures, 3.. replay
        la = a.step(x, y, lr=0.1)
        lb = fast(x, y)
        assert torch.allclose(la, lb, rtol=1e-5, atol=1e-6)
    m.flush()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-4, atol=1e-6)


def test_parallel_patterns(device):
    from mpi4jax_b200.parallel import (alltoall_reshard, average_gradients, broadcast_parameters,
                                       cartesian_neighbors, ring_shift)

    p = torch.full((3,), float(rank), device=device, requires_grad=True)
    broadcast_parameters([p], root=0, comm=comm)
    assert torch.equal(p.detach(), torch.zeros(3, device=device))
    p.grad = torch.full((3,), float(rank + 1), device=device)
    average_gradients([p], comm=comm)
    assert torch.allclose(p.grad, torch.full((3,), sum(range(1, size + 1)) / size, device=device))
    x = torch.full((4,), float(rank), device=device)
    assert torch.equal(ring_shift(x, 1, comm=comm), torch.full((4,), float((rank - 1) % size), device=device))
    # Ulysses reshard: (seq_local, heads) sequence-sharded -> (seq, heads_local) head-sharded
    seq_local, heads = 3, 2 * size
    t = (torch.arange(seq_local * heads, dtype=torch.float32).reshape(seq_local, heads)
         + 1000 * rank).to(device)
    out = alltoall_reshard(t, scatter_dim=1, gather_dim=0, comm=comm)
    assert out.shape == (seq_local * size, heads // size)
    for q in range(size):
        blk = out[q * seq_local:(q + 1) * seq_local].cpu()
        exp = (torch.arange(seq_local * heads, dtype=torch.float32).reshape(seq_local, heads)
               + 1000 * q)[:, rank * 2:(rank + 1) * 2]
        assert torch.equal(blk, exp)
    nb = cartesian_neighbors(0, 2, 4)
    assert nb == {"south": None, "north": 4, "west": 3, "east": 1}
