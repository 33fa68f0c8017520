"""The collective kernels synchronise their CTAs with each other and with the peer GPUs' CTAs
(csrc/b2_device.cuh: b2_barrier_all, the block-paired cross-GPU barrier).  That is only safe if every
CTA of a launch can be resident at the same time whatever else the GPU is running: the launch grids
are capped at the kernel's co-resident count (csrc/b2_collectives.cu: coresident_blocks = occupancy x SM
count), so a collective that is launched while other kernels own the SMs starts late but never deadlocks.

The reference has no counterpart (its collectives are MPI calls issued from a host callback,
mpi_xla_bridge_cuda.cpp); round 1's review asked for this property to be tested, not assumed."""

import pytest
import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()


def _occupy(device, rounds):
    """A stream of long GEMMs that fill every SM (8192^3 bf16, ~0.6 ms each), on a side stream."""
    side = torch.cuda.Stream(device=device)
    a = torch.randn(8192, 8192, device=device, dtype=torch.bfloat16)
    b = torch.randn(8192, 8192, device=device, dtype=torch.bfloat16)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        for _ in range(rounds):
            a = (a @ b).clamp_(-1, 1)
    return side, a


@pytest.mark.gpu
@pytest.mark.parametrize("nbytes", [1 << 10, 1 << 20, 64 << 20])
def test_collectives_complete_while_other_kernels_own_the_sms(nbytes):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    device = comm.device
    n = nbytes // 4
    k = min(n, 1024)
    x = torch.full((n,), float(rank + 1), device=device)
    blocks = torch.arange(size * 8, device=device, dtype=torch.float32).reshape(size, 8) + 100 * rank
    torch.cuda.synchronize(device)
    side, keep = _occupy(device, 24)            # ~15 ms of GEMMs in flight while the collectives launch
    for _ in range(3):
        total = m.allreduce(x, MPI.SUM, comm=comm)
        gathered = m.allgather(x[:k], comm=comm)
        swapped = m.alltoall(blocks, comm=comm)
        m.barrier(comm=comm)
    got = m.bcast(total[:16].clone(), 0, comm=comm)
    m.flush()
    torch.cuda.current_stream(device).wait_stream(side)
    torch.cuda.synchronize(device)
    want = float(size * (size + 1) // 2)
    assert torch.equal(total, torch.full((n,), want, device=device))
    assert torch.equal(got, torch.full((16,), want, device=device))
    assert torch.equal(gathered, torch.stack([torch.full((k,), float(q + 1), device=device) for q in range(size)]))
    for q in range(size):
        assert torch.equal(swapped[q], torch.arange(8, device=device, dtype=torch.float32) + rank * 8 + 100 * q)
    assert torch.isfinite(keep.float()).all()


@pytest.mark.gpu
def test_launch_grids_fit_the_gpu():
    """Every collective launch is capped at its kernel's co-resident CTA count (occupancy x SMs, the bound
    the barrier needs), itself bounded by the communicator's max_blocks: never more than two 512-thread
    CTAs per SM."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    nc = comm._native_comm()
    sms = torch.cuda.get_device_properties(comm.device).multi_processor_count
    assert nc.get_option("sm_count") == sms
    assert 1 <= nc.max_blocks <= 2 * sms
