"""ctypes binding of libucsa_nerf.so, generated from include/ucsa_nerf.h.

The header is the single source of truth: every ``UCSA_API`` declaration is parsed and bound with matching
ctypes argument types, so the Python side cannot drift from the C ABI.  There is no fallback: if the shared
library is missing or a symbol cannot be resolved, importing the operators raises.
"""
from __future__ import annotations

import ctypes
import os
import re
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(PKG_DIR), "include", "ucsa_nerf.h")
LIB_PATH = os.path.join(PKG_DIR, "csrc", "libucsa_nerf.so")

GRID_LEVELS = 16


class GridDesc(ctypes.Structure):
    """Mirror of ``ucsa_grid_desc``."""

    _fields_ = [
        ("scale", ctypes.c_float * GRID_LEVELS),
        ("res", ctypes.c_uint32 * GRID_LEVELS),
        ("entries", ctypes.c_uint32 * GRID_LEVELS),
        ("offset", ctypes.c_uint32 * GRID_LEVELS),
        ("hashed", ctypes.c_uint32 * GRID_LEVELS),
        ("total_entries", ctypes.c_uint32),
    ]


class UcsaError(RuntimeError):
    pass


_SCALARS = {
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "uint32_t": ctypes.c_uint32,
    "uint64_t": ctypes.c_uint64,
    "int32_t": ctypes.c_int32,
}


def _ctype_of(decl: str):
    decl = decl.strip()
    if decl == "void":
        return None
    if "*" in decl:
        if "ucsa_grid_desc" in decl:
            return ctypes.POINTER(GridDesc)
        return ctypes.c_void_p
    base = decl.replace("const", "").split()[0]
    return _SCALARS[base]


def parse_header(path: str = HEADER):
    """-> {name: (restype, [argtypes], [argnames])} for every UCSA_API declaration."""
    with open(path) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    out = {}
    for m in re.finditer(r"UCSA_API\s+(const char\*|int)\s+(ucsa_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        argtypes, argnames = [], []
        if args != "void":
            for a in args.split(","):
                a = a.strip()
                nm = re.search(r"(\w+)$", a).group(1)
                argtypes.append(_ctype_of(a[: -len(nm)]))
                argnames.append(nm)
        out[name] = (ctypes.c_char_p if ret.startswith("const char") else ctypes.c_int, argtypes, argnames)
    return out


_lock = threading.Lock()
_lib = None


def lib():
    """The loaded library with typed entry points; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise UcsaError(
                f"{LIB_PATH} is missing: build it with `python -m ucsa_neural_rendering_b200.build` "
                "(there is no CPU or PyTorch fallback for the rendering path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes, _) in parse_header().items():
            try:
                fn = getattr(handle, name)
            except AttributeError as exc:
                raise UcsaError(f"libucsa_nerf.so does not export {name}") from exc
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.ucsa_abi_version() != 1:
            raise UcsaError("libucsa_nerf.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().ucsa_last_error_string()
        raise UcsaError(f"{what or 'ucsa call'} failed ({rc}): {msg.decode() if msg else ''}")
