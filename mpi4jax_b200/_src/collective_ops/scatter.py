"""scatter -- rank q receives ``x_root[q]``.

Reference: /root/reference/mpi4jax/_src/collective_ops/scatter.py:44-91: on the root the
input has shape ``(nproc, *S)`` (error text below), elsewhere ``x`` is a template of shape S.
GPU: root stages, barrier, every rank pulls its block.
"""

from __future__ import annotations

import numpy as np

from ..comm import Comm
from ..utils import NOTSET, as_tensor, check_dtype, get_default_comm, raise_if_token_is_set
from ..validation import enforce_types
from . import _dispatch
from .bcast import _check_root


@enforce_types(root=(np.integer,), comm=(type(None), Comm))
def scatter(x, root, *, comm=None, token=NOTSET):
    """Perform a scatter operation.

    Returns:
        Tensor: this rank's block, shape ``x.shape[1:]`` on the root and ``x.shape`` elsewhere.
    """
    raise_if_token_is_set(token)
    if comm is None:
        comm = get_default_comm()
    x = as_tensor(x, comm)
    check_dtype(x)
    _check_root(root, comm, "Scatter")
    if comm.Get_rank() == root:
        if x.dim() == 0 or x.shape[0] != comm.Get_size():
            raise ValueError("Scatter input must have shape (nproc, ...)")
        out_shape = tuple(x.shape[1:])
    else:
        out_shape = tuple(x.shape)
    return _dispatch.scatter(comm, x, int(root), out_shape, x.dtype)
