"""Programmatic dependent launch on/off for the shallow-water launch chain (1 GPU).
Local size 1024x2048 == one rank of the 8-GPU 4096^2 run.  Also checks that PDL does not
change a single bit of the state."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mpi4jax_b200 as m  # noqa: E402
from mpi4jax_b200._src import native  # noqa: E402
from mpi4jax_b200.models import ShallowWaterConfig, ShallowWaterModel  # noqa: E402

for nx, ny in ((1024, 2048), (2048, 2048), (4096, 4096)):
    states = {}
    for pdl in (0, 1, 0, 1):
        native.lib.b2_set_pdl(pdl)
        mod = ShallowWaterModel(ShallowWaterConfig.for_resolution(nx, ny), device="cuda")
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
        same = ""
        if pdl in states:
            pass
        states.setdefault(pdl, st)
        if 0 in states and 1 in states:
            same = " bitwise_equal_to_other_mode=%s" % all(torch.equal(a, b) for a, b in zip(states[0], states[1]))
        print(f"nx={nx} ny={ny} pdl={pdl}: {us:.1f} us/step ({1e6 / us:.0f} steps/s) finite={bool(torch.isfinite(st[0]).all())}{same}",
              flush=True)
        del mod, run
native.lib.b2_set_pdl(0)
