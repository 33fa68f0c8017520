"""Timing hygiene helpers (B200_PROFILING.md): CUDA events on the launching stream,
synchronise on both sides, max over ranks, L2 flush between iterations, clock sampling."""

from __future__ import annotations

import statistics
import subprocess
import threading
import time
from typing import Callable, Optional

import torch


def max_over_ranks(value: float, comm=None) -> float:
    """Max of a host float over all ranks (control plane; not inside timed regions)."""
    import torch.distributed as dist

    from .._src.comm import get_world

    comm = comm or get_world()
    if comm.Get_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=comm._group)
    return float(t.item())


_l2_buf: Optional[torch.Tensor] = None


def flush_l2(device=None, nbytes: int = 256 << 20) -> None:
    """Write a buffer larger than the 126 MB L2."""
    global _l2_buf
    if _l2_buf is None or _l2_buf.numel() < nbytes:
        _l2_buf = torch.empty(nbytes, dtype=torch.uint8, device=device or "cuda")
    _l2_buf.fill_(1)


def device_time_ms(fn: Callable[[], None], iters: int = 10, warmup: int = 3, flush: bool = False,
                   comm=None, reduce: str = "median") -> float:
    """Device time of ``fn`` in ms: per-iteration CUDA events, max over ranks of the
    median (or mean/min) over iterations."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush:
            flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if comm is not None and comm.Get_size() > 1:
            comm.Barrier()
        s.record()
        fn()
        e.record()
        e.synchronize()
        times.append(s.elapsed_time(e))
    val = {"median": statistics.median, "mean": statistics.fmean, "min": min}[reduce](times)
    return max_over_ranks(val, comm)


class ClockSampler:
    """Samples ``nvidia-smi`` clocks / throttle reasons in the background during a timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0, period_s: float = 0.2):
        self.gpu_index, self.period_s = gpu_index, period_s
        self.samples: list = []
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                     "-i", str(self.gpu_index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(self.period_s)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=10)

    def summary(self) -> dict:
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[3:7]) if v.lower() == "active"})
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons,
            "samples": len(self.samples),
        }
