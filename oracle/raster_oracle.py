"""ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/raster_ref.c header: PARITY UNPINNED for the third-party rasterizer).

numpy/ctypes front-end of the C restatement.  Imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def build():
    src = os.path.join(_HERE, "raster_ref.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.siu3r_oracle_rasterize.restype = C.c_int
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def rasterize(means, cov6, shs, opacities, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, H, W, sh_degree, bg=(0, 0, 0), debug=True):
    """means [G,3], cov6 [G,6], shs [G,M,3] (the layout render_cuda passes), opacities [G]; matrices as torch lays them out
    (row-major view_matrix / full_projection of cuda_splatting.py:74-77).  Returns a dict of numpy arrays."""
    means, cov6, shs, opacities = _f(means), _f(cov6), _f(shs), _f(opacities)
    vm, pm, cp, bgv = _f(viewmatrix).reshape(16), _f(projmatrix).reshape(16), _f(campos).reshape(3), _f(bg).reshape(3)
    G, M = means.shape[0], shs.shape[1]
    color = np.zeros((3, H, W), np.float32)
    depth = np.zeros((H, W), np.float32)
    opac = np.zeros((H, W), np.float32)
    radii = np.zeros(G, np.int32)
    ntouch = np.zeros(G, np.int32)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    tiles = np.zeros(G, np.uint32)
    offs = np.zeros(G, np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    nren = C.c_int64(0)
    cap = 1 << 16
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    while True:
        keys = np.zeros(cap, np.uint64)
        vals = np.zeros(cap, np.uint32)
        rc = lib().siu3r_oracle_rasterize(C.c_int(G), C.c_int(H), C.c_int(W), C.c_int(sh_degree), C.c_int(M), P(means), P(cov6), P(shs), P(opacities),
                                          P(vm), P(pm), P(cp), C.c_float(tan_fovx), C.c_float(tan_fovy), P(bgv), P(color), P(depth), P(opac),
                                          P(radii), P(ntouch), P(tiles), P(offs), P(keys), P(vals), C.c_int64(cap), P(ranges), C.byref(nren))
        if rc == -2:
            cap = int(nren.value) + 16
            continue
        assert rc == 0, rc
        break
    D = int(nren.value)
    return dict(color=color, depth=depth, opacity=opac, radii=radii, n_touched=ntouch, tiles=tiles, offsets=offs, keys=keys[:D], values=vals[:D],
                ranges=ranges, num_rendered=D)


def rope2d(tokens: np.ndarray, pos: np.ndarray, base: float = 100.0, fwd: float = 1.0, cpu_order: bool = False) -> np.ndarray:
    """tokens [B,N,H,D] float32 (copied), pos [B,N,2] int64.  cpu_order: angle formed as in curope.cpp:36 instead of kernels.cu:44-53."""
    t = np.ascontiguousarray(tokens, dtype=np.float32).copy()
    p = np.ascontiguousarray(pos, dtype=np.int64)
    B, N, H, D = t.shape
    fn = lib().siu3r_oracle_rope2d_cpu if cpu_order else lib().siu3r_oracle_rope2d
    fn(t.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), C.c_int(B), C.c_int(N), C.c_int(H), C.c_int(D), C.c_float(base), C.c_float(fwd))
    return t
