"""Choice of the GPU transport of a communicator: native NVLink kernels or host staging.

``MPI4JAX_B200_TRANSPORT`` = ``auto`` (default) | ``native`` | ``host``.  ``auto`` uses the native
transport whenever all ranks of the communicator sit in one NVLink domain (one host) and falls
back to host staging -- the reference's default GPU mode -- with a warning otherwise.  The
reference's own switch is honoured too: ``MPI4JAX_USE_CUDA_MPI=0`` forces host staging
(/root/reference/mpi4jax/_src/decorators.py:38-64).
"""

from __future__ import annotations

import os
import socket
import warnings

from ..decorators import _is_falsy


def requested() -> str:
    """The transport this process asks for: 'auto', 'native' or 'host'."""
    want = os.environ.get("MPI4JAX_B200_TRANSPORT", "auto").strip().lower() or "auto"
    if want not in ("auto", "native", "host"):
        raise ValueError(f"MPI4JAX_B200_TRANSPORT must be auto, native or host (got {want!r})")
    raw = os.environ.get("MPI4JAX_USE_CUDA_MPI")
    if want == "auto" and raw is not None and _is_falsy(raw):
        want = "host"
    return want


def _node_name() -> str:
    """Hostname = NVLink domain.  ``MPI4JAX_B200_FAKE_NODE_SIZE=k`` (testing only) pretends that every
    k consecutive world ranks form their own node, so the multi-node paths can run on one machine."""
    fake = os.environ.get("MPI4JAX_B200_FAKE_NODE_SIZE")
    if fake:
        return f"fake-node-{int(os.environ.get('RANK', '0')) // max(1, int(fake))}"
    return socket.gethostname()


def decide(wants, hosts):
    """Pure decision function (unit-tested): every rank's request + hostname -> (transport, reason)."""
    if "host" in wants:
        return "host", "host staging requested (MPI4JAX_B200_TRANSPORT=host / MPI4JAX_USE_CUDA_MPI=0)"
    one_host = len(set(hosts)) == 1
    if "native" in wants:
        if not one_host:
            raise RuntimeError("MPI4JAX_B200_TRANSPORT=native, but the ranks of this communicator are on "
                               f"{len(set(hosts))} hosts: peer-mapped HBM needs a single NVLink domain")
        return "native", ""
    if one_host:
        return "native", ""
    return "host", f"ranks span {len(set(hosts))} hosts"


def create(comm):
    """Collective over ``comm``: build the GPU side of the communicator."""
    from ..decorators import setup_cuda_mpi
    from .cuda import _all_gather_obj

    setup_cuda_mpi()
    info = _all_gather_obj(comm, (requested(), _node_name()))
    kind, reason = decide([w for w, _ in info], [h for _, h in info])
    if kind == "native":
        from .cuda import NativeComm

        return NativeComm(comm)
    from .host_staged import HostStagedComm

    if comm.rank == 0 and "requested" not in reason:
        warnings.warn(f"mpi4jax_b200: {reason}; CUDA tensors of this communicator are staged through host "
                      "memory (gloo) instead of moving over NVLink. Set MPI4JAX_B200_TRANSPORT=host to "
                      "silence this warning.")
    return HostStagedComm(comm, reason)
