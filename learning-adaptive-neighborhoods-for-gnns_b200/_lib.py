"""ctypes binding of the C-ABI in include/dggb.h.  No CPU fallback: if libdggb.so is missing or a
call fails, this raises.

Every entry point gets its ``argtypes`` / ``restype`` from the prototype in the header (parsed once at load
time), so a call with the wrong arity or a wrong scalar type raises ``ctypes.ArgumentError`` in Python
instead of corrupting the callee's stack."""
from __future__ import annotations

import ctypes
import os
import re

from .build import LIB_PATH

_HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "dggb.h")
_lib = None


class DggbError(RuntimeError):
    pass


_SCALARS = {
    "int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64,
    "uint32_t": ctypes.c_uint32, "float": ctypes.c_float, "double": ctypes.c_double,
    "long long": ctypes.c_longlong, "size_t": ctypes.c_size_t,
}


def _header_text():
    text = open(_HEADER).read()
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def prototypes():
    """{name: (restype, [argtypes])} for every function include/dggb.h declares."""
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(dggb_[a-z0-9_]+)\s*\(([^;{]*)\)\s*;", _header_text()):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        protos[name] = (_ctype(ret), [] if params in ("", "void") else [_ctype(a) for a in params.split(",")])
    return protos


def _ctype(decl: str):
    decl = " ".join(decl.replace("const", " ").split())
    if "*" in decl:
        base = decl.split("*")[0].strip()
        return ctypes.c_char_p if base == "char" else ctypes.c_void_p
    # "int32_t n" -> "int32_t"; "long long" stays; a bare return type has no name part
    toks = decl.split()
    for width in (2, 1):
        cand = " ".join(toks[:width])
        if cand in _SCALARS:
            return _SCALARS[cand]
    raise DggbError(f"dggb.h: unknown C type in {decl!r}")


def declared_symbols():
    """Every entry point include/dggb.h declares (used by the symbol-export test)."""
    return sorted(set(re.findall(r"\b(dggb_[a-z0-9_]+)\s*\(", _header_text())))


def lib():
    global _lib
    if _lib is None:
        path = os.environ.get("DGGB_LIB", LIB_PATH)      # an alternative build of the same ABI (A/B measurements)
        if not os.path.isfile(path):
            raise DggbError(
                f"{path} not built: run `python __graft_entry__.py build` (there is no CPU fallback)")
        L = ctypes.CDLL(path)
        for name, (restype, argtypes) in prototypes().items():
            fn = getattr(L, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
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
    return None if t is None else t.data_ptr()


def stream():
    import torch

    return torch.cuda.current_stream().cuda_stream


i32 = int
i64 = int
f32 = float
