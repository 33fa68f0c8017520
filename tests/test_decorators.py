"""Environment-flag parsing and the extension guards (`_src/decorators.py`); covers the scenarios of
/root/reference/tests/test_decorators.py plus the flags that only exist here."""

import pytest


def test_truthy_falsy():
    from mpi4jax_b200._src.decorators import _is_falsy, _is_truthy

    for v in ("true", "True", "1", "on", "ON"):
        assert _is_truthy(v) and not _is_falsy(v)
    for v in ("false", "False", "0", "off", "OFF"):
        assert _is_falsy(v) and not _is_truthy(v)
    assert not _is_truthy("maybe") and not _is_falsy("maybe")


def test_env_flag(monkeypatch):
    from mpi4jax_b200._src.decorators import env_flag, env_float, env_int

    monkeypatch.delenv("B2_TEST_FLAG", raising=False)
    assert env_flag("B2_TEST_FLAG", True) is True
    monkeypatch.setenv("B2_TEST_FLAG", "off")
    assert env_flag("B2_TEST_FLAG", True) is False
    monkeypatch.setenv("B2_TEST_FLAG", "banana")
    with pytest.raises(RuntimeError):
        env_flag("B2_TEST_FLAG", True)
    monkeypatch.setenv("B2_TEST_INT", "4096")
    assert env_int("B2_TEST_INT", 1) == 4096
    assert env_float("B2_TEST_MISSING", 2.5) == 2.5


def test_ensure_cuda_ext(monkeypatch):
    from mpi4jax_b200._src import native
    from mpi4jax_b200._src.decorators import ensure_cuda_ext

    monkeypatch.setattr(native, "HAS_CUDA_EXT", True)
    ensure_cuda_ext()
    monkeypatch.setattr(native, "HAS_CUDA_EXT", False)
    with pytest.raises(ImportError) as excinfo:
        ensure_cuda_ext()
    assert "native CUDA library could not be loaded" in str(excinfo.value)

    @ensure_cuda_ext
    def f():
        return 1

    with pytest.raises(ImportError):
        f()


def test_ensure_xpu_ext():
    from mpi4jax_b200._src.decorators import ensure_xpu_ext

    with pytest.raises(ImportError) as excinfo:
        ensure_xpu_ext()
    assert "no XPU/SYCL extension" in str(excinfo.value)


def test_use_cuda_mpi_note(monkeypatch):
    import warnings

    from mpi4jax_b200._src import decorators

    monkeypatch.setattr(decorators, "_cuda_mpi_note_done", False)
    monkeypatch.setenv("MPI4JAX_USE_CUDA_MPI", "0")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        decorators.setup_cuda_mpi()
    assert any("host-staged" in str(x.message) for x in w)
