"""Launch schedules of the shallow-water step on one GPU: agreement with the stand-alone kernels
(bitwise), bitwise reproducibility of each, time per step at three local sizes (CUDA-graph replays)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mpi4jax_b200 as m  # noqa: E402
from mpi4jax_b200.models import ShallowWaterConfig, ShallowWaterModel  # noqa: E402

sizes = ((1024, 2048), (2048, 2048), (4096, 4096))
if len(sys.argv) > 1:
    sizes = tuple(tuple(int(x) for x in a.split("x")) for a in sys.argv[1:])
for nx, ny in sizes:
    states = {}
    for pipe in ("standalone", "ca", "ca"):
        mod = ShallowWaterModel(ShallowWaterConfig.for_resolution(nx, ny), device="cuda", pipeline=pipe)
        mod.step(first_step=True)
        run = m.jit(lambda: mod.multistep(50, first_step=False), warmup=0)
        run(); run()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            run()
        e.record(); e.synchronize()
        us = s.elapsed_time(e) * 1e3 / 500
        st = [t.clone() for t in mod.state]
        note = ""
        if pipe in states:
            note = " bitwise_repro=%s" % all(torch.equal(a, b) for a, b in zip(states[pipe], st))
        elif "standalone" in states:
            ref = states["standalone"]
            note = " bitwise_vs_standalone=%s max_rel_diff=%.2e" % (
                all(torch.equal(a[1:-1, 1:-1], b[1:-1, 1:-1]) for a, b in zip(st, ref)),
                max(((a - b).abs().max() / (b.abs().max() + 1e-30)).item() for a, b in zip(st, ref)))
        states.setdefault(pipe, st)
        print(f"nx={nx} ny={ny} {pipe}: {us:.1f} us/step ({1e6 / us:.0f} steps/s) "
              f"finite={bool(torch.isfinite(st[0]).all())}{note}", flush=True)
        del mod, run
