"""Single-GPU micro-benchmark of the shallow-water step at several local sizes, fused vs
stand-alone halo exchange (local size 1024x2048 == one rank of the 8-GPU 4096^2 run)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mpi4jax_b200 as m  # noqa: E402
from mpi4jax_b200.models import ShallowWaterConfig, ShallowWaterModel  # noqa: E402

for nx, ny in ((1024, 2048), (2048, 2048), (4096, 4096)):
    for fused in (True, False):
        mod = ShallowWaterModel(ShallowWaterConfig.for_resolution(nx, ny), device="cuda", fused=fused)
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
        print(f"nx={nx} ny={ny} fused={fused}: {us:.1f} us/step  ({1e6 / us:.0f} steps/s)", flush=True)
        del mod, run
