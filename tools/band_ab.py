#!/usr/bin/env python
"""tools/band_ab.py LIB [LIB ...] -- A/B of builds of libvrt_cuda.so on BASELINE configs 5 and 4 (run on the GPU box): one
process per build, median render time of 5 frames after a warm-up, CRC-32 of the packed image (a pure optimisation must
leave it unchanged).  LIB = a path, or "default" for the in-tree build.  Variant builds: e.g.
    nvcc <the Makefile's flags> -DBAND_RING=8 -o simd-gaussian-ray-tracing_b200/csrc/_ab/libvrt_cuda_ring8.so vrt_cuda.cu"""
import ctypes
import json
import os
import subprocess
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one(lib):
    sys.path.insert(0, ROOT)
    import numpy as np

    import __graft_entry__ as ge

    pkg = ge.load_package()
    if lib != "default":
        h = ctypes.CDLL(os.path.join(ROOT, lib))
        for sym, (res, args) in pkg._ffi.CUDA_SYMBOLS.items():
            fn = getattr(h, sym)
            fn.restype, fn.argtypes = res, args
        pkg._ffi._cuda = h
    V = pkg.vrt
    r = V.Renderer(0)
    out = {"lib": lib}
    for name, scene in (("config5", pkg.scenes.config5()), ("config4", pkg.scenes.config4())):
        W = 4096
        cam, origin = V.camera_t.app(W, W)
        r.set_gaussians(scene)
        f = r.frame(cam.view_matrix, origin, W, W, (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND, (256, 256), 6.0)
        ms = []
        for k in range(6):
            img, _, st = r.frame_render(f, True, False)
            if k:
                ms.append((st["ms_render"], st["ms_tile"]))
        out[name] = {"ms_render": round(float(np.median([m[0] for m in ms])), 3), "ms_tile": round(float(np.median([m[1] for m in ms])), 3), "exec": st["terms_executed"],
                     "sat": st["terms_saturated"], "term": st["terms_terminated"], "launches": st["n_launches"], "crc": zlib.crc32(img.tobytes())}
    r.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--one":
        one(sys.argv[2])
    else:
        for lib in sys.argv[1:] or ["default"]:
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one", lib], check=False)
