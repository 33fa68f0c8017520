"""gather -- the root receives every rank's ``x``: root ``S -> (nproc, *S)``.

Reference: /root/reference/mpi4jax/_src/collective_ops/gather.py:44-87: non-root ranks get
their *input* back.  GPU: stage -> barrier -> root pulls into the stacked layout.
"""

from __future__ import annotations

import numpy as np

from ..comm import Comm
from ..utils import NOTSET, as_tensor, check_dtype, get_default_comm, raise_if_token_is_set
from ..validation import enforce_types
from . import _dispatch
from .bcast import _check_root


@enforce_types(root=(np.integer,), comm=(type(None), Comm))
def gather(x, root, *, comm=None, token=NOTSET):
    """Perform a gather operation.

    Returns:
        Tensor: on the root ``(nproc, *x.shape)``; on every other rank the input ``x``.
    """
    raise_if_token_is_set(token)
    if comm is None:
        comm = get_default_comm()
    x = as_tensor(x, comm)
    check_dtype(x)
    _check_root(root, comm, "Gather")
    res = _dispatch.gather(comm, x, int(root))
    return res if comm.Get_rank() == root else x
