"""recv -- blocking-semantics point-to-point receive.

Reference: /root/reference/mpi4jax/_src/collective_ops/recv.py:47-74: ``x`` is only a
shape/dtype template; ``source`` defaults to ANY_SOURCE, ``tag`` to ANY_TAG; an optional
``status`` object is filled.  GPU: a kernel on the current stream waits for the message in
this rank's inbox ring and copies it out (csrc/b2_p2p.cu); ANY_SOURCE is resolved on the
device; the Status is read lazily.

Extension: differentiable -- the adjoint of a receive is a send of the cotangent back to
the source (requires a concrete ``source``).
"""

from __future__ import annotations

import numpy as np
import torch

from ..comm import ANY_SOURCE, ANY_TAG, PROC_NULL, Comm, Status
from ..utils import (NOTSET, as_tensor, check_dtype, check_rank, get_default_comm,
                     needs_autograd, raise_if_token_is_set)
from ..validation import enforce_types
from . import _dispatch


class _Recv(torch.autograd.Function):
    @staticmethod
    def forward(x, source, tag, comm, status):
        return _dispatch.recv(comm, x, source, tag, status)

    @staticmethod
    def setup_context(ctx, inputs, output):
        _, ctx.source, ctx.tag, ctx.comm, _ = inputs

    @staticmethod
    def backward(ctx, g):
        if ctx.source == ANY_SOURCE:
            raise RuntimeError("recv with source=ANY_SOURCE cannot be differentiated")
        comm, source, tag = ctx.comm, ctx.source, max(ctx.tag, 0)
        _dispatch.run_opaque(lambda t: _dispatch.send(comm, t, source, tag), g.contiguous())
        return None, None, None, None, None


@enforce_types(
    source=(np.integer,), tag=(np.integer,), comm=(type(None), Comm), status=(type(None), Status)
)
def recv(x, source=ANY_SOURCE, *, tag=ANY_TAG, comm=None, status=None, token=NOTSET):
    """Perform a recv (receive) operation.

    Arguments:
        x: template with the shape and dtype of the incoming message (not overwritten).
        source (int): sending rank (default: any).
        tag (int): tag to match (default: any).
        status: optional :class:`mpi4jax_b200.MPI.Status` to fill.

    Returns:
        Tensor: the received data.
    """
    raise_if_token_is_set(token)
    if comm is None:
        comm = get_default_comm()
    x = as_tensor(x, comm)
    check_dtype(x)
    check_rank(int(source), comm, "Recv", "source", allow_any=True)
    if int(source) == PROC_NULL:             # MPI: returns at once, buffer untouched
        if status is not None:
            status._set_proc_null()
        return x
    if not needs_autograd(x):
        return _dispatch.recv(comm, x, int(source), int(tag), status)
    return _Recv.apply(x, int(source), int(tag), comm, status)
