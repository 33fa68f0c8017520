def test_has_cuda_support():
    from mpi4jax_b200 import has_cuda_support

    assert isinstance(has_cuda_support(), bool)
    assert has_cuda_support(), "the native sm_100a library must be built (python __graft_entry__.py)"


def test_has_sycl_support():
    from mpi4jax_b200 import has_sycl_support

    assert has_sycl_support() is False


def test_native_abi_info():
    """Counterpart of the reference's MPI_ABI_INFO test (test_common.py:170-197)."""
    from mpi4jax_b200._src.native import NATIVE_ABI_INFO

    for key in ("library", "loaded", "version", "arch", "sizeof_status_record", "sizeof_halo_desc"):
        assert key in NATIVE_ABI_INFO
    assert NATIVE_ABI_INFO["arch"] == "sm_100a"
    assert NATIVE_ABI_INFO["sizeof_status_record"] == 24
