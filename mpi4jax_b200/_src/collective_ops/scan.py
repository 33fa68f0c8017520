"""scan -- inclusive prefix reduction over ranks: rank r gets ``op(x_0, ..., x_r)``.

Reference: /root/reference/mpi4jax/_src/collective_ops/scan.py:44-60, 113-114.
GPU: stage -> barrier -> rank r pulls the staged copies of ranks 0..r through the fused
reduce kernel (P <= 8 inside an NVLink domain, so O(P) direct reads beat a log-step chain).
"""

from __future__ import annotations

from ..comm import OP_TYPES, Comm, Op, as_op
from ..utils import NOTSET, as_tensor, check_dtype, get_default_comm, raise_if_token_is_set
from ..validation import enforce_types
from . import _dispatch


@enforce_types(op=OP_TYPES, comm=(type(None), Comm))
def scan(x, op, *, comm=None, token=NOTSET):
    """Perform a scan operation (inclusive prefix reduction over the ranks)."""
    raise_if_token_is_set(token)
    if comm is None:
        comm = get_default_comm()
    op = as_op(op)
    x = as_tensor(x, comm)
    check_dtype(x)
    return _dispatch.scan(comm, x, op.code)
