"""Packaging / version machinery (reference: versioneer, setup.py discovery, _latest_jax_version.txt)."""

import os
import re
import subprocess
import sys

import mpi4jax_b200 as m
from mpi4jax_b200._src import torch_compat
from mpi4jax_b200._version import _pep440

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_version_is_pep440():
    assert re.fullmatch(r"\d+\.\d+\.\d+(\+[0-9a-z.]+)?", m.__version__), m.__version__


def test_describe_to_pep440():
    assert _pep440("v0.1.0") == "0.1.0"
    assert _pep440("v0.1.0-0-gabc1234") == "0.1.0"
    assert _pep440("v0.1.0-5-gabc1234") == "0.1.0+5.gabc1234"
    assert _pep440("v0.1.0-5-gabc1234-dirty") == "0.1.0+5.gabc1234.dirty"
    assert _pep440("not-a-version") is None


def test_pinned_torch_version_file():
    path = os.path.join(REPO, "mpi4jax_b200", "_src", "_latest_torch_version.txt")
    txt = open(path).read().strip()
    assert txt.startswith("torch==")
    assert torch_compat.LATEST_TESTED_TORCH == txt.split("==")[1]
    assert torch_compat.versiontuple(torch_compat.MIN_TORCH) <= torch_compat.versiontuple(txt.split("==")[1])


def test_notset_repr_and_api_docs_are_current():
    assert repr(m._src.utils.NOTSET) == "NOTSET"
    before = open(os.path.join(REPO, "docs", "api.md")).read()
    for op in ("allreduce", "allgather", "alltoall", "barrier", "bcast", "gather", "reduce", "scan", "scatter",
               "send", "recv", "sendrecv", "has_cuda_support"):
        assert f"### `{op}(" in before, op


def test_setup_py_skip_native_build(tmp_path):
    """`MPI4JAX_B200_SKIP_NATIVE_BUILD=1 setup.py build_ext` is a no-op that succeeds without nvcc."""
    env = dict(os.environ, MPI4JAX_B200_SKIP_NATIVE_BUILD="1", PATH="/usr/bin:/bin")
    res = subprocess.run([sys.executable, "setup.py", "-q", "build_ext", "--build-lib", str(tmp_path),
                          "--build-temp", str(tmp_path)], cwd=REPO, env=env, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "skipping the native build" in res.stderr


def test_info_cli():
    res = subprocess.run([sys.executable, "-m", "mpi4jax_b200.info", "--json"], cwd=REPO, capture_output=True,
                         text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    import json

    info = json.loads(res.stdout)
    assert info["version"] == m.__version__ and info["native_loaded"] is True
    assert info["native_abi"]["abi_version"] >= 5 and info["requested_transport"] in ("auto", "native", "host")
    res = subprocess.run([sys.executable, "-m", "mpi4jax_b200.info"], cwd=REPO, capture_output=True, text=True,
                         timeout=120)
    assert res.returncode == 0 and "native core" in res.stdout
