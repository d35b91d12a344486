"""ctypes binding of the C-ABI in include/dggb.h.  No CPU fallback: if libdggb.so is missing or a
call fails, this raises."""
from __future__ import annotations

import ctypes
import os
import re

from .build import LIB_PATH

_HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "dggb.h")
_lib = None


class DggbError(RuntimeError):
    pass


def declared_symbols():
    """Every entry point include/dggb.h declares (used by the symbol-export test)."""
    text = open(_HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dggb_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise DggbError(
                f"{LIB_PATH} not built: run `python __graft_entry__.py build` (there is no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.dggb_error_string.restype = ctypes.c_char_p
        for name in declared_symbols():
            fn = getattr(_lib, name)
            if name in ("dggb_kernel_launches", "dggb_allpairs_workspace_bytes", "dggb_linear_act_workspace_bytes",
                        "dggb_gemm_tn_tc_workspace_bytes"):
                fn.restype = ctypes.c_longlong
            elif name != "dggb_error_string":
                fn.restype = ctypes.c_int
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        L = lib()
        msg = L.dggb_error_string(status).decode()
        if status == -5:
            msg += f" [cudaError {L.dggb_last_cuda_error()}]"
        raise DggbError(f"{what}: {msg}")


def p(t):
    """device pointer of a tensor (or NULL)."""
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


i32 = ctypes.c_int32
i64 = ctypes.c_int64
f32 = ctypes.c_float
