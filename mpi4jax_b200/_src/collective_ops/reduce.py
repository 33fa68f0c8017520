"""reduce -- the root receives the reduction of all ranks' ``x``.

Reference: /root/reference/mpi4jax/_src/collective_ops/reduce.py:45-71: non-root ranks get
their *input* back.  GPU: stage -> barrier -> root pulls P copies through the fused
reduce kernel (op + dtype cast in registers).
"""

from __future__ import annotations

import numpy as np

from ..comm import OP_TYPES, Comm, Op, as_op
from ..utils import NOTSET, as_tensor, check_dtype, get_default_comm, raise_if_token_is_set
from ..validation import enforce_types
from . import _dispatch
from .bcast import _check_root


@enforce_types(op=OP_TYPES, root=(np.integer,), comm=(type(None), Comm))
def reduce(x, op, root, *, comm=None, token=NOTSET):
    """Perform a reduce operation.

    Returns:
        Tensor: on the root the reduced data; on every other rank the input ``x``.
    """
    raise_if_token_is_set(token)
    if comm is None:
        comm = get_default_comm()
    op = as_op(op)
    x = as_tensor(x, comm)
    check_dtype(x)
    _check_root(root, comm, "Reduce")
    res = _dispatch.reduce(comm, x, op.code, int(root))
    return res if comm.Get_rank() == root else x
