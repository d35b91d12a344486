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
                               one, one, one, null) == -2
