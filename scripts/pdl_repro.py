"""Reproducibility probe: is the shallow-water state bitwise reproducible run-to-run, with and
without programmatic dependent launch?  Reports where the first differences appear."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mpi4jax_b200 as m  # noqa: E402
from mpi4jax_b200._src import native  # noqa: E402
from mpi4jax_b200.models import ShallowWaterConfig, ShallowWaterModel  # noqa: E402


def run(nx, ny, pdl, graph, nsteps):
    native.lib.b2_set_pdl(pdl)
    mod = ShallowWaterModel(ShallowWaterConfig.for_resolution(nx, ny), device="cuda")
    mod.step(first_step=True)
    if graph:
        f = m.jit(lambda: mod.multistep(10, first_step=False), warmup=0)
        for _ in range(nsteps // 10):
            f()
    else:
        mod.multistep(nsteps, first_step=False)
    torch.cuda.synchronize()
    return [t.clone() for t in mod.state]


def describe(a, b):
    out = []
    for name, x, y in zip("h u v dh du dv".split(), a, b):
        d = (x != y)
        if d.any():
            idx = d.nonzero()
            out.append(f"{name}: {int(d.sum())} cells, rows {int(idx[:, 0].min())}-{int(idx[:, 0].max())}, "
                       f"cols {int(idx[:, 1].min())}-{int(idx[:, 1].max())}, max|d|={float((x - y).abs().max()):.3e}")
    return "; ".join(out) or "IDENTICAL"


for nx, ny in ((1024, 2048), (512, 1024), (2048, 2048)):
    for graph in (False, True):
        for nsteps in (10, 100):
            base = run(nx, ny, 0, graph, nsteps)
            for pdl in (0, 1, 1):
                other = run(nx, ny, pdl, graph, nsteps)
                print(f"nx={nx} ny={ny} graph={graph} steps={nsteps} pdl0 vs pdl{pdl}: {describe(base, other)}", flush=True)
native.lib.b2_set_pdl(0)
