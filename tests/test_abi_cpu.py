"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/zkdl_b200.h
declares, host helpers agree with the oracle, and compute calls fail loudly without a GPU (no fallback)."""
import ctypes as C
import numpy as np
import pytest
from oracle import oracle as orc
from zkdl_b200 import capi


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = capi.declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/zkdl_b200.h but not exported"


def test_host_helpers_match_oracle():
    assert np.array_equal(capi.random_vec(7, 40), orc.random_vec(7, 40))
    for n in (0, 1, 2, 3, 1000, 1024, 1025, 784 * 1000):
        assert capi.lib().zkdl_ceil_log2(C.c_uint32(n)) == orc.ceil_log2(n)
    assert capi.lib().zkdl_partial_me_size(2 ** 19, 8, 2048) == 2048
    assert capi.lib().zkdl_partial_me_size(10, 2, 3) == 3
    assert capi.lib().zkdl_zkrelu_proof_size(1 << 19) == (3 * 24 + 1) + 32 + (3 * 23 + 1) + 16 + (3 * 19 + 2)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        capi.empty(4, 8)
    # raw ABI call without a device: must return an error code, not compute anything
    a = np.zeros((4, 8), np.uint32)
    rc = capi.lib().zkdl_fr_elementwise(0, a.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p), C.c_size_t(4), C.c_void_p(0))
    assert rc == 2 and b"" != capi.lib().zkdl_last_error()
