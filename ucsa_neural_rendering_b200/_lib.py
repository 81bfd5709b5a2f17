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
# UCSA_LIB: development override (A/B runs of two builds of the same ABI); the default is the in-tree build
LIB_PATH = os.environ.get("UCSA_LIB") or os.path.join(PKG_DIR, "csrc", "libucsa_nerf.so")

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
    "double": ctypes.c_double,
    "uint32_t": ctypes.c_uint32,
    "uint64_t": ctypes.c_uint64,
    "int32_t": ctypes.c_int32,
    "int64_t": ctypes.c_int64,
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

_DEBUG_SYNC = os.environ.get("UCSA_DEBUG_SYNC", "0") == "1"
HOST_ONLY = {"ucsa_abi_version", "ucsa_last_error_string", "ucsa_grid_desc_init"}
KERNELS_PER_CALL = {"ucsa_heads_fwd": 2, "ucsa_heads_bwd": 2}  # colour + semantic kernels; all others enqueue one


class LaunchStats:
    """Counts kernel launches (KERNELS_PER_CALL, default one per non-host entry point) and, for the entry points
    named in ``timed``, brackets each call with CUDA events on the current stream (used by bench.py only)."""

    def __init__(self):
        self.reset()
        self.timed = set()

    def reset(self):
        self.launches = 0
        self.by_name = {}
        self.events = {}

    def elapsed_ms(self, name):
        """-> list of per-launch durations; call after a device synchronize"""
        return [a.elapsed_time(b) for a, b in self.events.get(name, [])]


stats = LaunchStats()


class _Entry:
    __slots__ = ("name", "fn", "is_launch")

    def __init__(self, name, fn):
        self.name, self.fn, self.is_launch = name, fn, name not in HOST_ONLY

    def __call__(self, *args):
        if not self.is_launch:
            return self.fn(*args)
        stats.launches += KERNELS_PER_CALL.get(self.name, 1)
        stats.by_name[self.name] = stats.by_name.get(self.name, 0) + 1
        if _DEBUG_SYNC:  # UCSA_DEBUG_SYNC=1: name every launch and wait for it (locates a faulting kernel)
            import sys

            import torch

            sys.stderr.write(f"[ucsa] {self.name} ...")
            sys.stderr.flush()
            rc = self.fn(*args)
            torch.cuda.synchronize()
            sys.stderr.write(f" rc={rc}\n")
            sys.stderr.flush()
            return rc
        if self.name in stats.timed:
            import torch

            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = self.fn(*args)
            b.record()
            stats.events.setdefault(self.name, []).append((a, b))
            return rc
        return self.fn(*args)


class _Handle:
    pass


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
        cdll = ctypes.CDLL(LIB_PATH)
        handle = _Handle()
        for name, (restype, argtypes, _) in parse_header().items():
            try:
                fn = getattr(cdll, name)
            except AttributeError as exc:
                raise UcsaError(f"libucsa_nerf.so does not export {name}") from exc
            fn.restype = restype
            fn.argtypes = argtypes
            setattr(handle, name, _Entry(name, fn))
        if handle.ucsa_abi_version() != 1:
            raise UcsaError("libucsa_nerf.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().ucsa_last_error_string()
        raise UcsaError(f"{what or 'ucsa call'} failed ({rc}): {msg.decode() if msg else ''}")
