"""Workloads built on the primitives: the reference's shallow-water demo as a model class,
and a tensor/data-parallel MLP used for the autodiff benchmark configuration."""

from .shallow_water import (  # noqa: F401
    ModelState,
    ShallowWaterConfig,
    ShallowWaterModel,
    solve_shallow_water,
)
from .mlp import ParallelMLP  # noqa: F401
