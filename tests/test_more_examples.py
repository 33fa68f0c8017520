"""The small example programs under examples/ (run in-process at the current world size; the
multi-rank suite of tests/test_multirank.py runs them at 2 and 3 ranks as well)."""

import importlib.util
import os

import pytest

from mpi4jax_b200 import MPI

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
size = MPI.COMM_WORLD.Get_size()


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REPO, "examples", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_data_parallel_sgd_converges_and_replicas_agree():
    loss = _load("data_parallel_sgd").main(steps=80, verbose=False)
    assert loss < 1e-3


def test_distributed_transpose():
    assert _load("distributed_transpose").main(n=3, verbose=False)


def test_conjugate_gradient_with_sharded_operator():
    if 48 % size:
        pytest.skip("48 columns must divide evenly")
    assert _load("conjugate_gradient").main(verbose=False) < 1e-8
