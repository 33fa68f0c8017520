"""``python -m mpi4jax_b200.run -n N [--cpu] script.py [args...]`` -- mpirun stand-in.

The reference is launched with ``mpirun -n N python script.py`` (README.rst:83-89) and its
test-suite with ``mpirun -np 2 pytest .`` (docs/developers.rst:18-27).  This image has no
MPI launcher, and ``torchrun`` insists on a resolvable hostname, so this tiny launcher
starts N local ranks with RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR (default 127.0.0.1) /
MASTER_PORT set (``--nnodes / --node-rank / --master-addr / --master-port`` for several nodes),
prefixes nothing, waits for all of them and propagates the first non-zero exit code (killing the remaining ranks -- exactly the processes it started).

``-m module`` runs ``python -m module`` in every rank (``-m pytest tests/distributed``).
"""

from __future__ import annotations

import argparse
import os
import signal
import socket
import subprocess
import sys
import time


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def launch(nprocs: int, argv: list, *, cpu: bool = False, timeout: float | None = None,
           env_extra: dict | None = None, capture: bool = False, nnodes: int = 1, node_rank: int = 0,
           master_addr: str = "127.0.0.1", master_port: int | None = None, output_dir: str | None = None):
    """Start ``nprocs`` local ranks running ``python <argv...>``; returns (exit_code, outputs).

    Multi-node jobs run this launcher once per node with the same ``nnodes`` / ``master_addr`` /
    ``master_port`` and their own ``node_rank``; global rank = ``node_rank * nprocs + local rank``.
    Communicators that span nodes use the host-staged transport (``backends/transport.py``)."""
    if nnodes > 1 and master_port is None:
        raise ValueError("multi-node launches need an explicit --master-port (the same on every node)")
    port = master_port if master_port is not None else _free_port()
    procs = []
    for local in range(nprocs):
        rank = node_rank * nprocs + local
        env = dict(os.environ)
        env.update(
            RANK=str(rank), LOCAL_RANK=str(local), WORLD_SIZE=str(nnodes * nprocs),
            LOCAL_WORLD_SIZE=str(nprocs), MASTER_ADDR=master_addr, MASTER_PORT=str(port),
            GROUP_RANK=str(node_rank),
        )
        if cpu:
            env["MPI4JAX_B200_DEVICE"] = "cpu"
        if env_extra:
            env.update(env_extra)
        kw = dict(stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) if capture else {}
        if output_dir is not None:          # one log per rank (mpirun --output-filename)
            os.makedirs(output_dir, exist_ok=True)
            kw = dict(stdout=open(os.path.join(output_dir, f"rank{rank}.log"), "w"), stderr=subprocess.STDOUT)
        procs.append(subprocess.Popen([sys.executable, *argv], env=env, start_new_session=True, **kw))
    deadline = None if timeout is None else time.time() + timeout
    code = 0
    outputs = [""] * nprocs
    try:
        pending = set(range(nprocs))
        while pending:
            for r in list(pending):
                rc = procs[r].poll()
                if rc is None:
                    continue
                pending.discard(r)
                if rc != 0 and code == 0:
                    code = rc
                    # one rank failed: give the others a moment, then stop them
                    t_end = time.time() + 5
                    while time.time() < t_end and any(procs[q].poll() is None for q in pending):
                        time.sleep(0.1)
                    for q in pending:
                        if procs[q].poll() is None:
                            _kill(procs[q])
            if deadline is not None and time.time() > deadline:
                code = code or 124
                for q in pending:
                    _kill(procs[q])
                break
            time.sleep(0.05)
    finally:
        for r, p in enumerate(procs):
            if p.poll() is None:
                _kill(p)
            if capture:
                try:
                    outputs[r] = p.communicate(timeout=10)[0] or ""
                except Exception:
                    outputs[r] = ""
    return code, outputs


def _kill(p: subprocess.Popen) -> None:
    try:
        os.killpg(p.pid, signal.SIGKILL)      # the session we created for exactly this rank
    except (ProcessLookupError, PermissionError):
        pass


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m mpi4jax_b200.run", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-n", "--np", type=int, default=2, dest="nprocs", help="number of ranks")
    ap.add_argument("--cpu", action="store_true", help="force the CPU (gloo) backend")
    ap.add_argument("--timeout", type=float, default=None, help="kill the job after this many seconds")
    ap.add_argument("--nnodes", type=int, default=1, help="number of nodes (run the launcher once per node)")
    ap.add_argument("--node-rank", type=int, default=0, help="index of this node, 0 .. nnodes-1")
    ap.add_argument("--master-addr", default="127.0.0.1", help="address of node 0 (rendezvous)")
    ap.add_argument("--master-port", type=int, default=None, help="rendezvous port (required for nnodes > 1)")
    ap.add_argument("--output-dir", default=None, help="write every rank's stdout + stderr to <dir>/rank<r>.log")
    ap.add_argument("-m", dest="module", default=None, help="run a module (python -m ...) in every rank")
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    ns = ap.parse_args(argv)
    rest = ns.rest[1:] if ns.rest[:1] == ["--"] else ns.rest
    cmd = (["-m", ns.module] if ns.module else []) + rest
    if not cmd:
        ap.error("nothing to run")
    code, _ = launch(ns.nprocs, cmd, cpu=ns.cpu, timeout=ns.timeout, nnodes=ns.nnodes, node_rank=ns.node_rank,
                     master_addr=ns.master_addr, master_port=ns.master_port, output_dir=ns.output_dir)
    return code


if __name__ == "__main__":
    sys.exit(main())
