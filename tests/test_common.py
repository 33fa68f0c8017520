"""Failure / lifecycle / observability (scenario parity with /root/reference/tests/collective_ops/
test_common.py: subprocess harness, abort-on-error, deadlock-on-exit, debug logging)."""

import os
import re
import subprocess
import sys
from textwrap import dedent

import pytest
import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

comm = MPI.COMM_WORLD
rank = comm.Get_rank()
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_in_subprocess(code, test_file, timeout=120, device="cpu", extra_env=None):
    """Runs the given code in a fresh single-rank interpreter (scrubbed environment so that
    the child does not inherit RANK/WORLD_SIZE)."""
    test_file.write_text(code)
    env = dict(HOME=os.getenv("HOME", ""), PATH=os.getenv("PATH", ""), PYTHONPATH=REPO,
               LD_LIBRARY_PATH=os.getenv("LD_LIBRARY_PATH", ""), MPI4JAX_B200_DEVICE=device)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, str(test_file)], capture_output=True, timeout=timeout,
                          text=True, env=env)


@pytest.mark.skipif(rank > 0, reason="Runs only on rank 0")
def test_abort_on_error(tmp_path, device):
    script = dedent("""
        import torch
        import mpi4jax_b200 as m
        from mpi4jax_b200 import MPI

        comm = MPI.COMM_WORLD
        assert comm.Get_size() == 1
        x = torch.ones(10, device=comm.device)
        m.jit(lambda t: m.send(t, dest=100, comm=comm))(x)      # send to a non-existing rank
        m.jit(lambda t: m.send(t, dest=100, comm=comm))(x)
        m.flush()
    """)
    proc = run_in_subprocess(script, tmp_path / "abort.py", device=device.type)
    assert proc.returncode != 0
    assert "r0 | MPI_Send returned error code" in proc.stderr


@pytest.mark.skipif(rank > 0, reason="Runs only on rank 0")
def test_deadlock_on_exit(tmp_path, device):
    """Process exit with communication still in flight must neither hang nor crash."""
    script = dedent("""
        import torch
        import mpi4jax_b200 as m
        from mpi4jax_b200 import MPI

        comm = MPI.COMM_WORLD
        assert comm.Get_size() == 1
        x = torch.ones(10, device=comm.device)
        f = m.jit(lambda t: m.sendrecv(sendbuf=t, recvbuf=t, source=0, dest=0, comm=comm))
        for _ in range(3):
            f(x)
    """)
    proc = run_in_subprocess(script, tmp_path / "deadlock_on_exit.py", device=device.type)
    assert proc.returncode == 0, proc.stderr


def test_debug_logging(capsys, device):
    from mpi4jax_b200._src.native import set_logging

    arr = torch.ones((3, 2), device=device)
    m.allreduce(arr, op=MPI.SUM)      # make sure lazy initialisation is done
    m.flush()
    capsys.readouterr()
    set_logging(True)
    try:
        res = m.allreduce(arr, op=MPI.SUM)
        m.flush()
    finally:
        set_logging(False)
    captured = capsys.readouterr().out
    start_msg, end_msg, _ = captured.split("\n")
    assert re.match(rf"r{rank} \| \w{{8}} \| MPI_Allreduce( \(\w+\))? with {arr.numel()} items", start_msg)
    assert re.match(
        rf"r{rank} \| \w{{8}} \| MPI_Allreduce( \(\w+\))? done with code 0 \(\d\.\d{{2}}e[+-]?\d+s\)",
        end_msg)
    res = m.allreduce(arr, op=MPI.SUM)
    m.flush()
    assert not capsys.readouterr().out
    assert torch.equal(res, arr * comm.Get_size())


def test_set_logging_from_envvar(monkeypatch):
    import importlib

    from mpi4jax_b200._src import native

    monkeypatch.setenv("MPI4JAX_B200_DEBUG", "1")
    importlib.reload(native)
    assert native.get_logging()
    monkeypatch.setenv("MPI4JAX_B200_DEBUG", "0")
    importlib.reload(native)
    assert not native.get_logging()
    native.set_logging(True)
    assert native.get_logging()
    native.set_logging(False)
    monkeypatch.delenv("MPI4JAX_B200_DEBUG")
    monkeypatch.setenv("MPI4JAX_DEBUG", "1")      # the reference's variable is honoured too
    importlib.reload(native)
    assert native.get_logging()
    native.set_logging(False)


def test_comm_clone_and_split(device):
    size = comm.Get_size()
    c2 = comm.Clone()
    assert c2.Get_size() == size and c2.Get_rank() == rank and c2 != comm
    x = torch.ones(4, device=device)
    assert torch.equal(m.allreduce(x, MPI.SUM, comm=c2), x * size)
    sub = comm.Split(color=rank % 2, key=rank)
    assert sub.Get_size() == len(range(rank % 2, size, 2))
    assert torch.equal(m.allreduce(x, MPI.SUM, comm=sub), x * sub.Get_size())
    sub.Free()
    c2.Free()


def test_status_object():
    st = MPI.Status()
    st._set(3, 7, 40, itemsize=4)
    assert (st.Get_source(), st.Get_tag(), st.Get_count(), st.Get_count(MPI.BYTE)) == (3, 7, 10, 40)
    assert st.source == 3 and st.tag == 7


@pytest.mark.skipif(rank > 0, reason="Runs only on rank 0")
def test_nvtx_and_poison_debug_modes(tmp_path, device):
    """MPI4JAX_B200_NVTX wraps every native op in an NVTX range, MPI4JAX_B200_POISON fills fresh
    staging memory with 0xFF; both must leave results unchanged."""
    script = dedent("""
        import torch
        import mpi4jax_b200 as m
        from mpi4jax_b200 import MPI
        from mpi4jax_b200._src.backends import cuda as backend

        assert hasattr(backend.NativeComm.allreduce, "__wrapped__")      # NVTX wrapper installed
        assert backend._POISON
        comm = MPI.COMM_WORLD
        x = torch.arange(1000, dtype=torch.float32, device=comm.device)
        y = m.allreduce(x, MPI.SUM, comm=comm)
        z = m.allgather(x, comm=comm)
        big = m.allreduce(torch.ones(1 << 22, device=comm.device), MPI.SUM, comm=comm)
        m.flush()
        assert torch.equal(y, x) and torch.equal(z[0], x) and bool((big == 1).all())
        print("ok")
    """)
    proc = run_in_subprocess(script, tmp_path / "debug_modes.py", device=device.type,
                             extra_env={"MPI4JAX_B200_NVTX": "1", "MPI4JAX_B200_POISON": "1"})
    assert proc.returncode == 0, proc.stderr
    assert "ok" in proc.stdout
