"""Tiny driver for profiling: a few eager shallow-water steps on one GPU (4096x4096)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mpi4jax_b200.models import ShallowWaterConfig, ShallowWaterModel  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
model = ShallowWaterModel(ShallowWaterConfig.for_resolution(n, n), device="cuda")
model.step(first_step=True)
torch.cuda.synchronize()
model.multistep(steps)
torch.cuda.synchronize()
print("done", model.h.mean().item())
