"""barrier -- all ranks rendezvous.

Reference: /root/reference/mpi4jax/_src/collective_ops/barrier.py:42-57 (no inputs, no
outputs; vmap-able :100-102).  GPU: a device-side flag barrier enqueued on the current
stream (csrc/b2_collectives.cu, b2_k_barrier) -- it orders the *streams* of all ranks and
does not block the host; call ``mpi4jax_b200.flush()`` for a host-level rendezvous (the
reference needs ``jax.effects_barrier()`` for the same reason,
tests/collective_ops/test_barrier.py:38-39).
"""

from __future__ import annotations

from ..comm import Comm
from ..utils import NOTSET, get_default_comm, raise_if_token_is_set
from ..validation import enforce_types
from . import _dispatch


@enforce_types(comm=(type(None), Comm))
def barrier(*, comm=None, token=NOTSET):
    """Perform a barrier operation."""
    raise_if_token_is_set(token)
    if comm is None:
        comm = get_default_comm()
    _dispatch.barrier(comm)
