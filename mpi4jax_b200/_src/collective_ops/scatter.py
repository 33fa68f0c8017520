"""scatter -- rank q receives ``x_root[q]``.

Reference: /root/reference/mpi4jax/_src/collective_ops/scatter.py:44-91: on the root the
input has shape ``(nproc, *S)`` (error text below), elsewhere ``x`` is a template of shape S.
GPU: root stages, barrier, every rank pulls its block.
Extension: differentiable -- the adjoint of a scatter is a gather of the cotangents to the root
(non-root inputs are templates and get a zero gradient).
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


class _Scatter(torch.autograd.Function):
    @staticmethod
    def forward(x, root, comm, out_shape):
        return _dispatch.scatter(comm, x, root, out_shape, x.dtype)

    @staticmethod
    def setup_context(ctx, inputs, output):
        _, ctx.root, ctx.comm, _ = inputs

    @staticmethod
    def backward(ctx, g):
        comm, root = ctx.comm, ctx.root
        stacked = _dispatch.run_opaque(lambda t: _dispatch.gather(comm, t, root), g.contiguous())
        if comm.rank == root:
            return stacked, None, None, None
        return torch.zeros_like(g), None, None, None


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
    if needs_autograd(x):
        return _Scatter.apply(x, int(root), comm, out_shape)
    return _dispatch.scatter(comm, x, int(root), out_shape, x.dtype)
