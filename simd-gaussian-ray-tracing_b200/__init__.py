"""B200-native (sm_100a) render path of the `vrt` Gaussian ray tracer.

csrc/vrt_cuda.cu  CUDA kernels + the C ABI of include/vrt_cuda.h   -> csrc/libvrt_cuda.so
csrc/vrt_host.cpp CPU-side scene / camera / PNG helpers             -> csrc/libvrt_host.so
vrt.py            host-side mirror of the reference's vrt interface (camera_t, tile_gaussians, render_image, ...)
scenes.py         the callers' scene construction (grid, OBJ, synthetic)
bands.py          multi-GPU row bands over torch.distributed

The directory name carries a hyphen (it is the reference's repository name); import it through
`__graft_entry__.load_package()` which registers it as module `vrt_b200`.
"""
from . import _ffi, scenes, vrt  # noqa: F401
from .vrt import *  # noqa: F401,F403
