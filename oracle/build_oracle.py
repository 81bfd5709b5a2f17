"""Build recipes of the checker side (TEST INFRASTRUCTURE ONLY; see oracle/__init__.py).

* ``build_c_oracle``  gcc -> oracle/_build/libraymarch_oracle.so from oracle/raymarch_ref.c (the C restatement).
* ``build_reference`` when /root/reference is mounted (build container only): compiles the reference's own
  ``raymarching.cu`` + ``bindings.cpp`` *where they lie* into oracle/_ref/ as a torch extension (the sources need
  ATen, so this is the one recipe that uses torch.utils.cpp_extension; ``-std=c++17`` replaces the reference's
  ``-std=c++14`` (nr4seg/nerf/raymarching/backend.py:9,16), which torch >= 2.1 headers reject).  Nothing is copied.
  The GPU box has no /root/reference: it only loads the prebuilt oracle/_ref/_raymarching_ref.so.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
REF_DIR = os.path.join(HERE, "_ref")
C_LIB = os.path.join(BUILD_DIR, "libraymarch_oracle.so")
REF_SRC = "/root/reference/nr4seg/nerf/raymarching/src"
REF_LIB = os.path.join(REF_DIR, "_raymarching_ref.so")


def build_c_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "raymarch_ref.c")
    if not force and os.path.exists(C_LIB) and os.path.getmtime(C_LIB) >= os.path.getmtime(src):
        return C_LIB
    os.makedirs(BUILD_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", C_LIB, src, "-lm"]
    subprocess.run(cmd, check=True)
    return C_LIB


def build_reference(force: bool = False):
    """-> path of the built extension, or None when the reference sources are not mounted."""
    if not os.path.isdir(REF_SRC):
        return REF_LIB if os.path.exists(REF_LIB) else None
    if os.path.exists(REF_LIB) and not force:
        return REF_LIB
    os.makedirs(REF_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    load(name="_raymarching_ref", sources=[os.path.join(REF_SRC, "raymarching.cu"), os.path.join(REF_SRC, "bindings.cpp")],
         extra_cflags=["-O3", "-std=c++17"],
         extra_cuda_cflags=["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                            "-U__CUDA_NO_HALF2_OPERATORS__"],
         build_directory=REF_DIR, is_python_module=False, verbose=False)
    for junk in os.listdir(REF_DIR):
        if junk.endswith((".o", ".ninja", ".ninja_deps", ".ninja_log")) or junk == "build.ninja":
            os.remove(os.path.join(REF_DIR, junk))
    return REF_LIB if os.path.exists(REF_LIB) else None


def load_reference():
    """Import the prebuilt reference extension (GPU box or build container); None when unavailable."""
    if not os.path.exists(REF_LIB):
        return None
    import importlib.util

    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location("_raymarching_ref", REF_LIB)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_all():
    build_c_oracle()
    try:
        build_reference()
    except Exception as exc:  # the reference build is a bonus cross-check, never a requirement
        print(f"[oracle] reference extension not built: {exc}")


if __name__ == "__main__":
    build_all()
    print(C_LIB, os.path.exists(C_LIB), REF_LIB, os.path.exists(REF_LIB))
