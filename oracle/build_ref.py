"""ORACLE / TEST INFRASTRUCTURE: builds oracle/_ref/curope_ref.so = the reference's OWN curope.cpp (rope_2d -> rope_2d_cpu,
/root/reference/src/models/croco/curope/curope.cpp:11-65), compiled with g++ from where it lies (nothing is copied into this repository)
together with oracle/curope_cuda_stub.cpp.  It is the CPU form of the native entry point our fused RoPE replaces and is used by
tests/test_oracle_cpu.py to validate oracle/raster_ref.c: siu3r_oracle_rope2d.  External requirement: the torch headers / libraries of
this image (the same on the GPU box).  No-op when /root/reference is absent (GPU box: the prebuilt file travels with the snapshot).

The rest of the reference's native code is not buildable here: kernels.cu fails against torch 2.11 (kernels.cu:101) and
the two rasterizers (diff-gaussian-rasterization-w-pose, gsplat) are un-vendored dependencies whose sources are absent.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/models/croco/curope/curope.cpp"
OUT = os.path.join(HERE, "_ref", "curope_ref.so")


def build_ref(verbose: bool = True) -> str | None:
    if not os.path.exists(SRC):
        return OUT if os.path.exists(OUT) else None
    stub = os.path.join(HERE, "curope_cuda_stub.cpp")
    if os.path.exists(OUT) and os.path.getmtime(OUT) > max(os.path.getmtime(SRC), os.path.getmtime(stub)):
        return OUT
    import torch
    from torch.utils import cpp_extension as CE
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    inc = [f"-I{p}" for p in CE.include_paths()] + [f"-I{sysconfig.get_paths()['include']}"]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DTORCH_EXTENSION_NAME=curope_ref", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", *inc, SRC, stub, "-o", OUT,
           f"-L{libdir}", f"-Wl,-rpath,{libdir}", "-ltorch_python", "-ltorch", "-ltorch_cpu", "-lc10"]
    if verbose:
        print("[oracle.build_ref]", " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return OUT


def load_ref():
    """-> the compiled reference module (has rope_2d) or None when it was never built."""
    if not os.path.exists(OUT):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("curope_ref", OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build_ref())
