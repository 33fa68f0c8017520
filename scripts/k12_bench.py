"""Round-2 starter: fused flux+tendency path (k12) vs the stand-alone kernels on one GPU --
agreement, bitwise reproducibility of each, time per step at three local sizes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mpi4jax_b200 as m  # noqa: E402
from mpi4jax_b200.models import ShallowWaterConfig, ShallowWaterModel  # noqa: E402

for nx, ny in ((1024, 2048), (2048, 2048), (4096, 4096)):
    states = {}
    for k12 in (0, 1, 1, 2, 2):
        mod = ShallowWaterModel(ShallowWaterConfig.for_resolution(nx, ny), device="cuda", k12=k12)
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
        if k12 in states:
            note = " bitwise_repro=%s" % all(torch.equal(a, b) for a, b in zip(states[k12], st))
        elif 0 in states:
            ref = states[0]
            note = " max_rel_diff_vs_standalone=%.2e" % max(
                ((a - b).abs().max() / (b.abs().max() + 1e-30)).item() for a, b in zip(st, ref))
        states.setdefault(k12, st)
        print(f"nx={nx} ny={ny} k12={k12}: {us:.1f} us/step ({1e6 / us:.0f} steps/s) "
              f"finite={bool(torch.isfinite(st[0]).all())}{note}", flush=True)
        del mod, run
