"""Transport selection and the host-staged fallback (the reference's default GPU mode:
mpi_xla_bridge_cuda.cpp stages through host memory unless MPI4JAX_USE_CUDA_MPI=1)."""

import pytest
import torch

import mpi4jax_b200 as m
from mpi4jax_b200 import MPI
from mpi4jax_b200._src.backends import host_staged, transport
from mpi4jax_b200._src.native import codes

comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()


def test_decide():
    assert transport.decide(["auto", "auto"], ["a", "a"]) == ("native", "")
    kind, reason = transport.decide(["auto", "auto"], ["a", "b"])
    assert kind == "host" and "2 hosts" in reason
    assert transport.decide(["auto", "host"], ["a", "a"])[0] == "host"
    assert transport.decide(["native", "native"], ["a", "a"])[0] == "native"
    with pytest.raises(RuntimeError, match="single NVLink domain"):
        transport.decide(["native", "auto"], ["a", "b"])


def test_requested_env(monkeypatch):
    monkeypatch.delenv("MPI4JAX_B200_TRANSPORT", raising=False)
    monkeypatch.delenv("MPI4JAX_USE_CUDA_MPI", raising=False)
    assert transport.requested() == "auto"
    monkeypatch.setenv("MPI4JAX_USE_CUDA_MPI", "0")
    assert transport.requested() == "host"            # the reference's switch is honoured
    monkeypatch.setenv("MPI4JAX_USE_CUDA_MPI", "1")
    assert transport.requested() == "auto"
    monkeypatch.setenv("MPI4JAX_B200_TRANSPORT", "native")
    monkeypatch.setenv("MPI4JAX_USE_CUDA_MPI", "0")
    assert transport.requested() == "native"
    monkeypatch.setenv("MPI4JAX_B200_TRANSPORT", "bogus")
    with pytest.raises(ValueError):
        transport.requested()


def test_host_staged_ops_match_reference_semantics(device, monkeypatch):
    """All 12 ops through the staging class (device-agnostic, so the CPU suite covers the logic;
    on a GPU box the tensors really travel D2H -> gloo -> H2D)."""
    monkeypatch.setattr(host_staged, "ACTIVE", host_staged.ACTIVE)      # restore the flag afterwards
    hs = host_staged.HostStagedComm(comm.Clone(), "unit test")
    c = hs.comm
    x = torch.arange(6, dtype=torch.float32, device=device) + rank
    tot = sum(range(size))
    assert torch.equal(hs.allreduce(x, codes.SUM), torch.arange(6, device=device) * size + tot)
    assert hs.allreduce(x, codes.SUM).device == x.device
    g = hs.allgather(x)
    assert g.shape == (size, 6) and torch.equal(g[rank], x)
    a = torch.arange(size * 2, dtype=torch.float32, device=device).reshape(size, 2) + 100 * rank
    at = hs.alltoall(a)
    assert torch.equal(at[:, 0], torch.arange(size, device=device) * 100.0 + 2 * rank)
    b = hs.bcast(x, 0)
    assert torch.equal(b, torch.arange(6, dtype=torch.float32, device=device))
    r = hs.reduce(x, codes.SUM, 0)
    if rank == 0:
        assert torch.equal(r, torch.arange(6, device=device) * size + tot)
    s = hs.scan(x, codes.SUM)
    assert torch.equal(s, torch.arange(6, device=device) * (rank + 1) + sum(range(rank + 1)))
    ga = hs.gather(x, 0)
    if rank == 0:
        assert ga.shape == (size, 6)
    sc = hs.scatter(a if rank == 0 else a[0], 0, (2,), a.dtype)
    assert torch.equal(sc, torch.tensor([2.0 * rank, 2.0 * rank + 1], device=device))
    status = MPI.Status()
    out = hs.sendrecv(x, torch.empty_like(x), (rank - 1) % size, (rank + 1) % size, 3, 3, status)
    assert torch.equal(out, torch.arange(6, dtype=torch.float32, device=device) + (rank - 1) % size)
    assert status.Get_source() == (rank - 1) % size and status.Get_tag() == 3
    hs.send(x, rank, 7)
    assert torch.equal(hs.recv(torch.empty_like(x), rank, 7, None), x)
    hs.barrier()
    with pytest.raises(MPI.MPIError, match="native NVLink transport"):
        hs.halo_exchange()
    c.Free()


def test_jit_runs_eagerly_when_host_staged(monkeypatch):
    from mpi4jax_b200._src import jit as jitmod

    monkeypatch.setattr(host_staged, "ACTIVE", True)
    assert jitmod._host_staged()
    calls = []
    f = m.jit(lambda t: (calls.append(1), t + 1)[1])
    x = torch.zeros(2)
    for _ in range(3):
        assert torch.equal(f(x), x + 1)
    assert len(calls) == 3            # never replayed from a graph


def test_transport_property_on_cpu():
    if comm.device.type == "cpu":
        assert comm.transport == "cpu"
    else:
        assert comm.transport in ("native", "host")


@pytest.mark.skipif(rank > 0, reason="Runs only on rank 0")
def test_host_transport_end_to_end(tmp_path, device):
    """A whole process with MPI4JAX_B200_TRANSPORT=host: public ops, jit, the model's backend choice."""
    from textwrap import dedent

    from .test_common import run_in_subprocess

    script = dedent("""
        import torch
        import mpi4jax_b200 as m
        from mpi4jax_b200 import MPI
        from mpi4jax_b200.models import ShallowWaterConfig, ShallowWaterModel

        comm = MPI.COMM_WORLD
        dev = comm.device
        assert comm.transport == ("host" if dev.type == "cuda" else "cpu"), comm.transport
        x = torch.arange(8, dtype=torch.float32, device=dev)
        y = m.jit(lambda t: m.allreduce(t * 2, MPI.SUM, comm=comm))(x)
        for _ in range(3):
            y = m.jit(lambda t: m.allreduce(t * 2, MPI.SUM, comm=comm))(x)
        assert torch.equal(y, 2 * x) and y.device == x.device
        z = m.sendrecv(x, torch.empty_like(x), 0, 0, comm=comm)
        assert torch.equal(z, x)
        g = torch.func.grad(lambda t: m.allreduce(t * t, MPI.SUM, comm=comm).sum())(x)
        assert torch.equal(g, 2 * x)
        model = ShallowWaterModel(ShallowWaterConfig(nx=24, ny=12), comm=comm, device=dev)
        assert model.backend == "ops"
        model.multistep(3)
        assert torch.isfinite(model.h).all()
        m.barrier(comm=comm)
        m.flush()
        print("ok", comm.transport)
    """)
    proc = run_in_subprocess(script, tmp_path / "host_transport.py", device=device.type,
                             extra_env={"MPI4JAX_B200_TRANSPORT": "host"})
    assert proc.returncode == 0, proc.stderr[-3000:]
    assert "ok" in proc.stdout


_inside_job = size > 1 or "MPI4JAX_B200_NESTED" in __import__("os").environ

_FALLBACK_JOB = """
import warnings
import torch
import mpi4jax_b200 as m
from mpi4jax_b200 import MPI

comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    kind = comm.transport
want = "host" if comm.device.type == "cuda" else "cpu"
assert kind == want, (kind, want)
if rank == 0 and want == "host":
    assert any("ranks span 2 hosts" in str(x.message) for x in w), [str(x.message) for x in w]
x = torch.arange(5, dtype=torch.float32, device=comm.device) + rank
assert torch.equal(m.allreduce(x, MPI.SUM, comm=comm), 2 * torch.arange(5, device=comm.device) + 1.0)
other = 1 - rank
y = m.sendrecv(x, torch.empty_like(x), other, other, comm=comm)
assert torch.equal(y, torch.arange(5, dtype=torch.float32, device=comm.device) + other)
if rank == 0:
    m.send(x, 1, tag=4)
else:
    assert torch.equal(m.recv(torch.empty_like(x), 0, tag=4), x - 1)
sub = comm.Split(rank, 0)                 # one rank per fake node: single-host again -> native on GPU
assert sub.transport == ("native" if comm.device.type == "cuda" else "cpu")
assert torch.equal(m.allreduce(x, MPI.SUM, comm=sub), x)
m.barrier(comm=comm)
m.flush()
print("fallback ok", rank, kind)
"""


def _run_fallback_job(tmp_path, cpu):
    import os

    from mpi4jax_b200.run import launch

    script = tmp_path / "fallback_job.py"
    script.write_text(_FALLBACK_JOB)
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code, outs = launch(2, [str(script)], cpu=cpu, timeout=240, capture=True,
                        env_extra={"MPI4JAX_B200_FAKE_NODE_SIZE": "1", "MPI4JAX_B200_NESTED": "1",
                                   "PYTHONPATH": repo})
    assert code == 0, "\n".join(o[-3000:] for o in outs)
    assert all("fallback ok" in o for o in outs)


@pytest.mark.skipif(_inside_job, reason="already inside a multi-rank job")
def test_multi_node_job_on_cpu(tmp_path):
    """Two fake nodes with CPU tensors: nothing changes (gloo is the transport anyway)."""
    _run_fallback_job(tmp_path, cpu=True)


@pytest.mark.gpu
@pytest.mark.skipif(_inside_job, reason="already inside a multi-rank job")
def test_multi_node_job_falls_back_to_host_staging(tmp_path):
    """Two GPU ranks that pretend to sit on different nodes: the world communicator picks the
    host-staged transport (with a warning), a per-node sub-communicator stays native."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    _run_fallback_job(tmp_path, cpu=False)


def test_launcher_multi_node_rank_layout(tmp_path):
    """--nnodes/--node-rank: global rank = node_rank * nprocs + local rank (no rendezvous needed)."""
    import json

    from mpi4jax_b200.run import launch

    script = tmp_path / "env.py"
    script.write_text("import os, json; print(json.dumps({k: os.environ[k] for k in "
                      "('RANK','LOCAL_RANK','WORLD_SIZE','LOCAL_WORLD_SIZE','MASTER_ADDR','MASTER_PORT')}))")
    code, outs = launch(2, [str(script)], capture=True, nnodes=3, node_rank=2, master_addr="10.0.0.1",
                        master_port=29999, timeout=60)
    assert code == 0
    envs = sorted((json.loads(o.strip().splitlines()[-1]) for o in outs), key=lambda e: int(e["RANK"]))
    assert [e["RANK"] for e in envs] == ["4", "5"] and [e["LOCAL_RANK"] for e in envs] == ["0", "1"]
    assert all(e["WORLD_SIZE"] == "6" and e["MASTER_ADDR"] == "10.0.0.1" and e["MASTER_PORT"] == "29999"
               for e in envs)
    with pytest.raises(ValueError, match="master-port"):
        launch(1, [str(script)], nnodes=2)
