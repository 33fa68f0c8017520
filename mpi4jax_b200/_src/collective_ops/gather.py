"""gather -- the root receives every rank's ``x``: root ``S -> (nproc, *S)``.

Reference: /root/reference/mpi4jax/_src/collective_ops/gather.py:44-87: non-root ranks get
their *input* back.  GPU: stage -> barrier -> root pulls into the stacked layout.
Extension: differentiable -- the adjoint of a gather is a scatter of the root's cotangent (a
non-root rank's result IS its input, so its own cotangent passes through as well).
"""

from __future__ import annotations

import numpy as np

import torch

from ..comm import Comm
from ..utils import (NOTSET, as_tensor, check_dtype, get_default_comm, needs_autograd,
                     raise_if_token_is_set)
from ..validation import enforce_types
from . import _dispatch
from .bcast import _check_root


class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(x, root, comm):
        res = _dispatch.gather(comm, x, root)
        return res if comm.rank == root else x.view_as(x)

    @staticmethod
    def setup_context(ctx, inputs, output):
        x, ctx.root, ctx.comm = inputs
        ctx.meta = (tuple(x.shape), x.dtype)

    @staticmethod
    def backward(ctx, g):
        comm, root = ctx.comm, ctx.root
        shape, dtype = ctx.meta
        if comm.rank == root:
            mine = _dispatch.run_opaque(lambda t: _dispatch.scatter(comm, t, root, shape, dtype), g.contiguous())
            return mine, None, None
        template = torch.empty(shape, dtype=dtype, device=g.device)
        return _dispatch.scatter(comm, template, root, shape, dtype) + g, None, None


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
    if needs_autograd(x):
        return _Gather.apply(x, int(root), comm)
    res = _dispatch.gather(comm, x, int(root))
    return res if comm.Get_rank() == root else x
