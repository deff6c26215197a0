"""Scene construction for the hot path's callers (thin numpy wrappers over libvrt_host.so, include/vrt_host.h).

grid(dim)                  `-g <dim>` of src/volumetric-ray-tracer/main.cpp:196-204
img_error_grid()           the 16x16 grid of tests/img-error.cpp:18-26
transmittance_test()       the three Gaussians of tests/transmittance.cpp:9
read_obj(path)             `-f <file>`: read_from_obj, src/vrt/gaussians-from-file.cpp:7-44
synthetic(n, seed, lo, hi) frustum-filling random scene of BASELINE configs 4/5
"""
import ctypes

import numpy as np

from . import _ffi


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def grid(dim):
    out = np.zeros((dim * dim, 10), np.float32)
    n = _ffi.host_lib().vrt_host_scene_grid(dim, _ptr(out))
    if n != dim * dim:
        raise ValueError(f"grid dim {dim} out of range (1..255, u8 loop indices in main.cpp:196-198)")
    return out


def grid_ex(dim, sigma, magnitude):
    out = np.zeros((dim * dim, 10), np.float32)
    n = _ffi.host_lib().vrt_host_scene_grid_ex(dim, sigma, magnitude, _ptr(out))
    if n != dim * dim:
        raise ValueError(f"grid dim {dim} out of range")
    return out


def img_error_grid():
    return grid_ex(16, 0.25, 3.0)


def transmittance_test():
    out = np.zeros((3, 10), np.float32)
    _ffi.host_lib().vrt_host_scene_transmittance_test(_ptr(out))
    return out


def read_obj(path):
    lib = _ffi.host_lib()
    n = lib.vrt_host_read_obj(path.encode(), None, 0)
    if n == 2**64 - 1:
        raise OSError(f"cannot read {path}")
    out = np.zeros((n, 10), np.float32)
    lib.vrt_host_read_obj(path.encode(), _ptr(out), n)
    return out


def synthetic(n, seed, log10_sigma_lo, log10_sigma_hi):
    out = np.zeros((n, 10), np.float32)
    _ffi.host_lib().vrt_host_scene_synthetic(n, seed, log10_sigma_lo, log10_sigma_hi, _ptr(out))
    return out


# BASELINE.json configs 4 and 5
def config4():
    return synthetic(100_000, 42, -2.1, -1.5)


def config5():
    return synthetic(1_000_000, 43, -2.6, -2.0)
