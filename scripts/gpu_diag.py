"""Incremental GPU bring-up diagnostic (run under `python -m mpi4jax_b200.run -n N`).

Exercises every native layer in order, printing one line per step, so a failure (or hang,
bounded by the device watchdog) pinpoints the layer.  Also prints first latency/bandwidth
numbers next to NCCL.  Output: gpurun_out/diag_rank<r>.log
"""

import os
import sys
import time
import traceback

os.environ.setdefault("MPI4JAX_B200_TIMEOUT", "20")
os.environ.setdefault("MPI4JAX_B200_ABORT_ON_ERROR", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import mpi4jax_b200 as m  # noqa: E402
from mpi4jax_b200 import MPI  # noqa: E402
from mpi4jax_b200._src import native  # noqa: E402
from mpi4jax_b200._src.native import codes  # noqa: E402

comm = MPI.COMM_WORLD
rank, size = comm.Get_rank(), comm.Get_size()
os.makedirs("gpurun_out", exist_ok=True)
LOG = open(f"gpurun_out/diag_rank{rank}.log", "w")


def say(*a):
    msg = " ".join(str(x) for x in a)
    LOG.write(msg + "\n")
    LOG.flush()
    if rank == 0:
        print(msg, flush=True)


def step(name):
    def deco(fn):
        t0 = time.time()
        try:
            out = fn()
            m.flush()
            say(f"[{name}] OK {time.time() - t0:.2f}s {out if out is not None else ''}")
        except Exception as exc:
            say(f"[{name}] FAIL {type(exc).__name__}: {exc}")
            say(traceback.format_exc())
        return fn
    return deco


dev = comm.device
say("device", dev, torch.cuda.get_device_name(dev), "world", size)


@step("native-init")
def _():
    try:
        nc = comm._native_comm()
    except Exception as exc:
        say("native init with VMM failed, retrying with cudaIpc:", exc)
        os.environ["MPI4JAX_B200_HEAP"] = "ipc"
        comm._native = None
        nc = comm._native_comm()
    return f"mode={nc.mode} nvls={nc.has_nvls} shared_gpu={nc.shared_gpu}"


@step("barrier")
def _():
    for _ in range(5):
        m.barrier(comm=comm)


def check_allreduce(n, dtype, algo):
    x = (torch.arange(n, device=dev) % 7 + rank).to(dtype)
    out = m.allreduce(x, MPI.SUM, comm=comm, algorithm=algo)
    exp = ((torch.arange(n, device=dev) % 7) * size + sum(range(size))).to(dtype)
    assert torch.equal(out, exp), (algo, n, dtype, (out.float() - exp.float()).abs().max().item())


for algo, sizes in (("ll", (1, 5, 1000, 8000)), ("oneshot", (1, 1000, 100_003, 1 << 20)),
                    ("twoshot", (1000, 100_003, 1 << 22)), ("nvls", (1 << 12, 100_004, 1 << 22))):
    for dt in (torch.float32, torch.bfloat16):
        @step(f"allreduce-{algo}-{str(dt).split('.')[-1]}")
        def _(algo=algo, sizes=sizes, dt=dt):
            if algo == "nvls" and not comm._native_comm().has_nvls:
                return "skipped (no multicast)"
            for n in sizes:
                check_allreduce(n, dt, algo)


@step("allreduce-auto-int-max")
def _():
    x = torch.arange(1000, device=dev, dtype=torch.int32) + rank
    assert torch.equal(m.allreduce(x, MPI.MAX, comm=comm), torch.arange(1000, device=dev, dtype=torch.int32) + size - 1)


@step("reduce-scan")
def _():
    x = torch.ones(1000, device=dev) * (rank + 1)
    r = m.reduce(x, MPI.SUM, 0, comm=comm)
    if rank == 0:
        assert r[0].item() == sum(range(1, size + 1))
    s = m.scan(x, MPI.SUM, comm=comm)
    assert s[0].item() == sum(range(1, rank + 2))


@step("allgather-alltoall")
def _():
    x = torch.ones(1000, device=dev) * rank
    g = m.allgather(x, comm=comm)
    assert torch.equal(g[:, 0].cpu(), torch.arange(size, dtype=torch.float32))
    a = torch.stack([torch.ones(333, device=dev) * (rank * 10 + q) for q in range(size)])
    r = m.alltoall(a, comm=comm)
    assert torch.equal(r[:, 0].cpu(), torch.tensor([q * 10 + rank for q in range(size)], dtype=torch.float32))
    big = torch.ones(size, 1 << 20, device=dev) * rank
    r = m.alltoall(big, comm=comm)
    assert torch.equal(r[:, 5].cpu(), torch.arange(size, dtype=torch.float32))


@step("bcast-gather-scatter")
def _():
    x = torch.arange(5000, device=dev, dtype=torch.float32) * (1 if rank == 0 else 0)
    b = m.bcast(x, 0, comm=comm)
    assert torch.equal(b, torch.arange(5000, device=dev, dtype=torch.float32))
    g = m.gather(torch.ones(100, device=dev) * rank, 0, comm=comm)
    if rank == 0:
        assert torch.equal(g[:, 0].cpu(), torch.arange(size, dtype=torch.float32))
    src = torch.stack([torch.ones(100) * q for q in range(size)]).to(dev) if rank == 0 else torch.empty(100, device=dev)
    s = m.scatter(src, 0, comm=comm)
    assert s[0].item() == rank


@step("p2p-sendrecv-ring")
def _():
    for n in (1, 1000, 70_001, 5_000_003):
        x = torch.arange(n, device=dev, dtype=torch.float32) + rank
        st = MPI.Status()
        r = m.sendrecv(x, x, source=(rank - 1) % size, dest=(rank + 1) % size, comm=comm, status=st)
        assert torch.equal(r, torch.arange(n, device=dev, dtype=torch.float32) + (rank - 1) % size), n
        assert st.Get_source() == (rank - 1) % size and st.Get_count() == n


@step("p2p-send-recv-anysource")
def _():
    if size < 2:
        return "skipped"
    x = torch.ones(100, device=dev) * rank
    if rank == 0:
        seen = set()
        for _ in range(size - 1):
            st = MPI.Status()
            r = m.recv(x, comm=comm, status=st)
            assert r[0].item() == st.Get_source()
            seen.add(st.Get_source())
        assert seen == set(range(1, size))
    else:
        m.send(x, 0, tag=rank, comm=comm)


@step("halo-and-swe-native-vs-ops")
def _():
    from mpi4jax_b200.models import ShallowWaterConfig, ShallowWaterModel
    cfg = ShallowWaterConfig(nx=128, ny=64)
    a = ShallowWaterModel(cfg, comm=comm, device=dev, backend="native")
    b = ShallowWaterModel(cfg, comm=comm, device=dev, backend="ops")
    err0 = max((x - y).abs().max().item() for x, y in zip(a.state, b.state))
    a.multistep(10)
    b.multistep(10)
    err = [(x - y).abs().max().item() for x, y in zip(a.state, b.state)]
    scale = [y.abs().max().item() for y in b.state]
    return f"init err {err0:.2e}; after 10 steps abs err {['%.2e' % e for e in err]} scale {['%.2e' % s for s in scale]}"


@step("jit-graph-allreduce")
def _():
    f = m.jit(lambda x: m.allreduce(x, MPI.SUM, comm=comm) * 2)
    x = torch.ones(4096, device=dev)
    for i in range(5):
        assert f(x + i)[0].item() == 2 * size * (1 + i)


@step("autograd-mlp")
def _():
    from mpi4jax_b200.models import ParallelMLP
    for mode in ("dp", "tp"):
        mlp = ParallelMLP(16, 32, 4, comm=comm, device=dev, mode=mode)
        torch.manual_seed(rank)
        l0 = mlp.step(torch.randn(8, 16, device=dev), torch.randn(8, 4, device=dev))
        assert torch.isfinite(l0)


# ---------------------------------------------------------------- first numbers
def timeit(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    comm.Barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) / iters * 1e3   # us


@step("perf-allreduce")
def _():
    import torch.distributed as dist
    nccl = None
    try:
        nccl = dist.new_group(backend="nccl")
    except Exception as exc:
        say("nccl group unavailable:", exc)
    rows = []
    has_nvls = comm._native_comm().has_nvls
    for nbytes in (1 << 10, 1 << 14, 1 << 17, 1 << 20, 1 << 23, 1 << 26, 1 << 28):
        x = torch.ones(nbytes // 4, device=dev)
        row = [f"{nbytes:>10d}B"]
        for algo in ("auto", "ll", "oneshot", "twoshot", "nvls"):
            if algo == "ll" and nbytes > (64 << 10):
                row.append("ll=-")
                continue
            if algo == "nvls" and not has_nvls:
                row.append("nvls=-")
                continue
            if algo == "oneshot" and nbytes > (1 << 26) and size > 2:
                row.append("oneshot=-")
                continue
            try:
                us = timeit(lambda: m.allreduce(x, MPI.SUM, comm=comm, algorithm=algo), iters=10 if nbytes > (1 << 24) else 30)
                bus = nbytes / us / 1e3 * 2 * (size - 1) / size
                row.append(f"{algo}={us:.1f}us/{bus:.0f}GB/s")
            except Exception as exc:
                row.append(f"{algo}=ERR({exc})")
        if nccl is not None:
            us = timeit(lambda: dist.all_reduce(x, group=nccl), iters=10 if nbytes > (1 << 24) else 30)
            row.append(f"nccl={us:.1f}us/{nbytes / us / 1e3 * 2 * (size - 1) / size:.0f}GB/s")
        rows.append(" ".join(row))
        say(rows[-1])
    return ""


@step("perf-p2p-halo")
def _():
    out = []
    for nbytes in (4096, 1 << 20, 1 << 26):
        x = torch.ones(nbytes // 4, device=dev)
        us = timeit(lambda: m.sendrecv(x, x, source=(rank - 1) % size, dest=(rank + 1) % size, comm=comm), iters=20)
        out.append(f"sendrecv {nbytes}B {us:.1f}us {nbytes / us / 1e3:.1f}GB/s")
    from mpi4jax_b200.models import ShallowWaterConfig, ShallowWaterModel
    mod = ShallowWaterModel(ShallowWaterConfig.for_resolution(4096, 4096), comm=comm, device=dev)
    us = timeit(lambda: mod.enforce_boundaries([mod.fe, mod.fn, mod.q, mod.ke], ["u", "v", "h", "h"]), iters=50)
    out.append(f"halo4 {us:.1f}us")
    run = m.jit(lambda: mod.multistep(20, first_step=False))
    run(); run()
    us = timeit(run, iters=5, warmup=2)
    out.append(f"swe 4096^2 graph: {us / 20:.1f}us/step = {20e6 / us:.0f} steps/s")
    us = timeit(lambda: mod.multistep(20, first_step=False), iters=3, warmup=1)
    out.append(f"swe eager: {us / 20:.1f}us/step")
    return "; ".join(out)


say("launches", native.launch_count())
say("DIAG DONE")
LOG.close()
