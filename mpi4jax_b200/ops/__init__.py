"""Functional op namespace: the 12 primitives plus the fused ops that go beyond the
reference's one-MPI-call-per-op model (fused multi-field halo exchange)."""

from .._src import (  # noqa: F401
    allgather, allreduce, alltoall, barrier, bcast, gather, recv, reduce, scan, scatter,
    send, send_with_grad, sendrecv,
)
from .halo import halo_exchange  # noqa: F401
from .linear import linear_allreduce  # noqa: F401
