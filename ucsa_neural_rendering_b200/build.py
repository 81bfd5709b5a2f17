"""In-tree build of libucsa_nerf.so (sm_100a) with plain nvcc.

    python -m ucsa_neural_rendering_b200.build [--force]

Objects and the shared library stay next to the sources (git-ignored) so that they travel with the tree.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import re
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(CSRC, "libucsa_nerf.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-DUCSA_BUILDING_LIBRARY",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libucsa_nerf.so cannot be built")
    return exe


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


_INCLUDE = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)
INCLUDE_DIR = os.path.join(os.path.dirname(PKG_DIR), "include")


def _closure(path: str, seen: set) -> None:
    """path and every file it includes with quotes (searched next to it, then in include/)."""
    if path in seen:
        return
    seen.add(path)
    with open(path, "r", errors="replace") as fh:
        text = fh.read()
    for name in _INCLUDE.findall(text):
        for base in (os.path.dirname(path), INCLUDE_DIR):
            cand = os.path.normpath(os.path.join(base, name))
            if os.path.exists(cand):
                _closure(cand, seen)
                break


def _source_stamp(src: str) -> str:
    """Hash of the flags, the source and its include closure: an object is rebuilt only when one of those changed."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps: set = set()
    _closure(os.path.join(CSRC, src), deps)
    for path in sorted(deps):
        h.update(path.encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    sources = _sources()
    stamps = {src: _source_stamp(src) for src in sources}

    def stamp_path(src):
        return os.path.join(OBJ_DIR, src[:-3] + ".stamp")

    def fresh(src):
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        if force or not os.path.exists(obj) or not os.path.exists(stamp_path(src)):
            return False
        with open(stamp_path(src)) as fh:
            return fh.read().strip() == stamps[src]

    stale = [src for src in sources if not fresh(src)]
    lib_stamp_file = os.path.join(OBJ_DIR, "stamp")
    lib_stamp = hashlib.sha256("".join(stamps[s] for s in sources).encode()).hexdigest()
    if not stale and os.path.exists(LIB_PATH) and os.path.exists(lib_stamp_file):
        with open(lib_stamp_file) as fh:
            if fh.read().strip() == lib_stamp:
                return LIB_PATH
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE_DIR, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
        with open(stamp_path(src), "w") as fh:
            fh.write(stamps[src])
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        list(pool.map(compile_one, stale))
    objs = [os.path.join(OBJ_DIR, src[:-3] + ".o") for src in sources]
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs, "-lcudart"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(lib_stamp_file, "w") as fh:
        fh.write(lib_stamp)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
