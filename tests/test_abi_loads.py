"""CPU: the C-ABI library builds, loads and exports every symbol include/dggb.h declares."""
import ctypes

import dgg_b200


def test_library_exports_every_declared_symbol():
    dgg_b200.build.build()
    L = dgg_b200.lib()
    names = dgg_b200.declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), n
    assert L.dggb_version() >= 1
    assert L.dggb_build_arch() == 1000
    assert L.dggb_error_string(0) == b"ok"
    assert b"shape" in L.dggb_error_string(-2)


def test_bad_arguments_are_rejected_without_touching_the_gpu():
    L = dgg_b200.lib()
    null = ctypes.c_void_p(0)
    i32 = ctypes.c_int32
    assert L.dggb_spmm_csr_fwd(null, null, null, i32(4), null, i32(8), null, null, null) == -1
    assert L.dggb_sym_normalize_fwd(null, null, null, i32(4), null, null, null) == -1
    one = ctypes.c_void_p(16)  # never dereferenced: shape check fires first
    assert L.dggb_dgg_edge_fwd(one, one, one, i32(4), i32(9), i32(6), one, one, one, one, null, i32(-1), one, one,
                               one, one, one, null, null) == -2
    # typed prototypes (argtypes from the header): a drifted call is an exception, not a corrupted stack
    import pytest
    with pytest.raises(ctypes.ArgumentError):
        L.dggb_sym_normalize_fwd(null, null, null, "four", null, null, null)
    with pytest.raises(TypeError):
        L.dggb_sym_normalize_fwd(null, null, null)
    # the fused edge-probability kernels validate their flag / extras combination before touching memory
    assert L.dggb_edge_mlp_fwd(one, one, 8, 64, 128, one, null, 0, null, null, null, one, one, one, 0.01, 1.0, 16, one,
                               null) == -3
    assert L.dggb_edge_mlp_fwd(one, one, 8, 62, 128, one, null, 0, null, null, null, one, one, one, 0.01, 1.0, 0, one,
                               null) == -2
