"""send -- blocking-semantics point-to-point send.

Reference: /root/reference/mpi4jax/_src/collective_ops/send.py:44-64 (returns nothing).
GPU: the message is pushed into the destination's inbox ring over NVLink by a kernel on
the current stream (csrc/b2_p2p.cu); the host never blocks.

Extension (BASELINE north star "VJP rules on send/recv"): ``send`` cannot carry a gradient
through a ``None`` result, so pipeline-style code uses :func:`send_with_grad`, which
returns a scalar token; back-propagating through the token *receives* the cotangent of
``x`` from ``dest`` (the adjoint of a send is a receive from the same peer).
"""

from __future__ import annotations

import numpy as np
import torch

from ..comm import PROC_NULL, Comm
from ..utils import (NOTSET, as_tensor, check_dtype, carries_grad, check_rank, get_default_comm,
                     raise_if_token_is_set)
from ..validation import enforce_types
from . import _dispatch


class _SendWithGrad(torch.autograd.Function):
    @staticmethod
    def forward(x, dest, tag, comm):
        _dispatch.send(comm, x, dest, tag)
        return x.new_zeros(())

    @staticmethod
    def setup_context(ctx, inputs, output):
        x, ctx.dest, ctx.tag, ctx.comm = inputs
        ctx.meta = (tuple(x.shape), x.dtype, x.device)      # not a tensor: x may be a torch.func wrapper

    @staticmethod
    def backward(ctx, _g):
        shape, dtype, device = ctx.meta
        comm, dest, tag = ctx.comm, ctx.dest, ctx.tag

        def receive(_token_cotangent):
            return _dispatch.recv(comm, torch.empty(shape, dtype=dtype, device=device), dest, tag, None)

        # (under torch.func.grad even a tensor created HERE is a transform wrapper without storage; the
        # backend, which takes raw pointers on the GPU path, has to run below the transform levels)
        return _dispatch.run_opaque(receive, _g), None, None, None


@enforce_types(dest=(np.integer,), tag=(np.integer,), comm=(type(None), Comm))
def send(x, dest, *, tag=0, comm=None, token=NOTSET):
    """Perform a send operation.

    Arguments:
        x: tensor, array or scalar to send.
        dest (int): rank of the destination process.
        tag (int): message tag (default 0).
    """
    raise_if_token_is_set(token)
    if comm is None:
        comm = get_default_comm()
    x = as_tensor(x, comm)
    check_dtype(x)
    check_rank(int(dest), comm, "Send", "destination")
    if int(dest) == PROC_NULL:
        return
    if carries_grad(x):
        # a None result cannot carry a gradient: silently cutting the graph here would drop it
        raise NotImplementedError(
            "send of a tensor that requires grad: use mpi4jax_b200.send_with_grad (its token's backward "
            "pass receives the cotangent), or send x.detach()")
    _dispatch.send(comm, x.detach(), int(dest), int(tag))


@enforce_types(dest=(np.integer,), tag=(np.integer,), comm=(type(None), Comm))
def send_with_grad(x, dest, *, tag=0, comm=None):
    """Differentiable send: returns a zero scalar whose backward pass receives the
    cotangent of ``x`` from ``dest`` (pair it with a differentiable :func:`recv` there)."""
    if comm is None:
        comm = get_default_comm()
    x = as_tensor(x, comm)
    check_dtype(x)
    check_rank(int(dest), comm, "Send", "destination")
    return _SendWithGrad.apply(x, int(dest), int(tag), comm)
