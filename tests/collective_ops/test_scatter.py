"""scatter: the root's ``(nproc, *S)`` input is dealt out, rank r receives block r
(reference semantics: /root/reference/mpi4jax/_src/collective_ops/scatter.py:44-90, 145-153;
non-root callers pass a template of the block's shape and dtype)."""

import pytest
import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI


comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()

ROOTS = sorted({0, size - 1})
BLOCKS = [((3, 2), torch.float32), ((5,), torch.int64), ((2, 1, 4), torch.float64), ((), torch.float32)]


def _deck(shape, dtype, device):
    """What the root scatters: block r is filled with 10 r + (0, 1, 2, ...)."""
    n = 1
    for s in shape:
        n *= s
    base = torch.arange(n, dtype=dtype, device=device).reshape(shape)
    return torch.stack([base + 10 * r for r in range(size)])


@pytest.mark.parametrize("root", ROOTS)
@pytest.mark.parametrize("shape, dtype", BLOCKS, ids=lambda v: str(v).replace("torch.", ""))
def test_every_rank_gets_its_block(device, root, shape, dtype):
    deck = _deck(shape, dtype, device)
    arg = deck if rank == root else torch.empty(shape, dtype=dtype, device=device)
    before = arg.clone()
    out = m.scatter(arg, root=root)
    assert out.shape == torch.Size(shape) and out.dtype == dtype and out.device == arg.device
    assert torch.equal(out, deck[rank])
    assert torch.equal(arg, before)                       # inputs are never written to


def test_replay_under_jit(device):
    deck = _deck((3, 2), torch.float32, device)
    arg = deck if rank == 0 else torch.empty((3, 2), device=device)
    dealt = m.jit(lambda t: m.scatter(t, root=0))
    for _ in range(3):
        assert torch.equal(dealt(arg), deck[rank])


def test_leading_axis_must_equal_the_communicator_size(device):
    if rank != 0:
        return
    for bad in (torch.ones((size + 1, 3, 2), device=device), torch.ones((size + 2,), device=device)):
        with pytest.raises(ValueError, match=r"Scatter input must have shape \(nproc, \.\.\.\)"):
            m.scatter(bad, root=0)
