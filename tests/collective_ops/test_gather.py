"""gather: the root receives every rank's array stacked along a new leading axis, the other ranks
get their input back (reference: /root/reference/mpi4jax/_src/collective_ops/gather.py:44-88,
140-150)."""

import pytest
import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI


comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()
ROOTS = sorted({0, size - 1})


def _piece(r, shape, dtype, device):
    n = 1
    for s in shape:
        n *= s
    return (torch.arange(n, device=device).reshape(shape) + 100 * r).to(dtype)


@pytest.mark.parametrize("root", ROOTS)
@pytest.mark.parametrize("shape, dtype", [((3, 2), torch.float32), ((4,), torch.float32), ((2, 3), torch.int32),
                                          ((1,), torch.bool)], ids=lambda v: str(v).replace("torch.", ""))
def test_root_collects_all_pieces_in_rank_order(device, root, shape, dtype):
    x = _piece(rank, shape, dtype, device)
    keep = x.clone()
    out = m.gather(x, root=root)
    if rank == root:
        assert out.shape == (size, *shape) and out.dtype == dtype
        for r in range(size):
            assert torch.equal(out[r], _piece(r, shape, dtype, device))
    else:
        assert torch.equal(out, x)
    assert torch.equal(x, keep)


def test_python_scalar(device):
    out = m.gather(rank, root=0)
    if rank == 0:
        assert out.cpu().tolist() == list(range(size))
    else:
        assert out.item() == rank


def test_non_contiguous_input(device):
    x = _piece(rank, (4, 6), torch.float32, device).t()[::2]          # strided view, shape (3, 4)
    out = m.gather(x, root=0)
    if rank == 0:
        for r in range(size):
            assert torch.equal(out[r], _piece(r, (4, 6), torch.float32, device).t()[::2])


def test_replay_under_jit(device):
    x = _piece(rank, (3, 2), torch.float32, device)
    collect = m.jit(lambda t: m.gather(t, root=0))
    for _ in range(3):
        out = collect(x)
        if rank == 0:
            assert torch.equal(out, torch.stack([_piece(r, (3, 2), torch.float32, device) for r in range(size)]))
        else:
            assert torch.equal(out, x)
