"""ctypes bindings of the two product libraries (no torch types cross this boundary).

libvrt_cuda.so  -- include/vrt_cuda.h, the sm_100a render path.  Loading it needs only the CUDA runtime;
                   creating a context needs a GPU.  There is NO fallback: if the library is missing the
                   import of `cuda_lib()` raises, it never routes to a CPU implementation.
libvrt_host.so  -- include/vrt_host.h, CPU-side scene / camera / PNG helpers.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")

c_f, c_d, c_i, c_u32, c_u64, vp = ctypes.c_float, ctypes.c_double, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p


class Frame(ctypes.Structure):
    """struct vrt_cuda_frame (include/vrt_cuda.h)."""

    _fields_ = [
        ("view", c_f * 16),
        ("origin", c_f * 4),
        ("width", c_u32),
        ("height", c_u32),
        ("tiles_x", c_u32),
        ("tiles_y", c_u32),
        ("flags", c_u32),
        ("bound_sigmas", c_f),
        ("row_begin", c_u32),
        ("row_end", c_u32),
    ]


class Stats(ctypes.Structure):
    """struct vrt_cuda_stats (include/vrt_cuda.h)."""

    _fields_ = [
        ("n_gaussians", c_u64),
        ("n_cells", c_u64),
        ("list_entries", c_u64),
        ("max_list", c_u32),
        ("n_launches", c_u32),
        ("terms_listed", c_d),
        ("terms_executed", c_d),
        ("ms_tile", c_f),
        ("ms_render", c_f),
        ("ms_total", c_f),
        ("slice", c_u32),
        ("terms_saturated", c_d),
        ("terms_terminated", c_d),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/vrt_cuda.h declares: name -> (restype, argtypes)
CUDA_SYMBOLS = {
    "vrt_cuda_abi_version": (c_i, []),
    "vrt_cuda_create": (c_i, [c_i, ctypes.POINTER(vp)]),
    "vrt_cuda_destroy": (None, [vp]),
    "vrt_cuda_last_error": (ctypes.c_char_p, [vp]),
    "vrt_cuda_set_gaussians": (c_i, [vp, vp, c_u64]),
    "vrt_cuda_set_gaussians_device": (c_i, [vp, vp, c_u64]),
    "vrt_cuda_tile": (c_i, [vp, ctypes.POINTER(Frame)]),
    "vrt_cuda_set_tile_lists": (c_i, [vp, ctypes.POINTER(Frame), vp, vp, c_u64]),
    "vrt_cuda_get_lists": (c_i, [vp, vp, c_u64, vp, c_u64, ctypes.POINTER(c_u64), ctypes.POINTER(c_u64)]),
    "vrt_cuda_render": (c_i, [vp, ctypes.POINTER(Frame), vp, vp, ctypes.POINTER(Stats)]),
    "vrt_cuda_render_device": (c_i, [vp, ctypes.POINTER(Frame), vp, vp, ctypes.POINTER(Stats)]),
    "vrt_cuda_frame_render": (c_i, [vp, ctypes.POINTER(Frame), vp, vp, ctypes.POINTER(Stats)]),
    "vrt_cuda_render_interruptible": (c_i, [vp, ctypes.POINTER(Frame), vp, vp, ctypes.POINTER(Stats), vp]),
    "vrt_cuda_abort": (c_i, [vp, c_i]),
    "vrt_cuda_set_host_pinning": (c_i, [vp, c_i]),
    "vrt_cuda_pin_buffer": (c_i, [vp, vp, c_u64]),
    "vrt_cuda_unpin_buffer": (c_i, [vp, vp]),
    "vrt_cuda_peer_image_create": (c_i, [vp, c_u64, ctypes.POINTER(vp), vp]),
    "vrt_cuda_peer_image_open": (c_i, [vp, vp, ctypes.POINTER(vp)]),
    "vrt_cuda_peer_image_close": (c_i, [vp, vp]),
    "vrt_cuda_row_costs": (c_i, [vp, vp, c_u32, ctypes.POINTER(c_u32), ctypes.POINTER(c_u32)]),
    "vrt_cuda_set_tuning": (c_i, [vp, c_i, c_i]),
    "vrt_cuda_set_band_tuning": (c_i, [vp, c_i]),
    "vrt_cuda_fp32_peak": (c_i, [vp, c_i, ctypes.POINTER(c_d)]),
    "vrt_cuda_term_peak": (c_i, [vp, c_i, c_i, ctypes.POINTER(c_d)]),
    "vrt_cuda_mix_peak": (c_i, [vp, c_i, c_i, c_i, ctypes.POINTER(c_d)]),
    "vrt_cuda_approx_table": (c_i, [vp, c_i, vp, vp, c_u64]),
    "vrt_cuda_approx_rate": (c_i, [vp, c_i, ctypes.POINTER(c_d)]),
    "vrt_cuda_set_slice": (c_i, [vp, c_i]),
    "vrt_cuda_auto_slice": (c_i, [vp, c_d, ctypes.POINTER(c_i)]),
    "vrt_cuda_sync": (c_i, [vp]),
    "vrt_cuda_stream": (c_u64, [vp]),
    "vrt_cuda_device": (c_i, [vp]),
}

HOST_SYMBOLS = {
    "vrt_host_scene_grid": (c_u64, [c_u32, vp]),
    "vrt_host_scene_grid_ex": (c_u64, [c_u32, c_f, c_f, vp]),
    "vrt_host_scene_transmittance_test": (c_u64, [vp]),
    "vrt_host_scene_synthetic": (c_u64, [c_u64, c_u64, c_f, c_f, vp]),
    "vrt_host_read_obj": (c_u64, [ctypes.c_char_p, vp, c_u64]),
    "vrt_host_view_matrix": (None, [vp, c_f, c_f, c_f, vp]),
    "vrt_host_app_camera": (None, [c_f, c_f, c_f, vp, vp]),
    "vrt_host_row_bands": (c_i, [vp, c_u32, c_u32, vp]),
    "vrt_host_write_png": (c_i, [ctypes.c_char_p, c_u32, c_u32, vp]),
}

_cuda = None
_host = None


def _load(name, symbols):
    path = os.path.join(CSRC, name)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `make -C {CSRC}` (or __graft_entry__.build()); there is no fallback path")
    lib = ctypes.CDLL(path)
    for sym, (res, args) in symbols.items():
        fn = getattr(lib, sym)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


def cuda_lib():
    global _cuda
    if _cuda is None:
        _cuda = _load("libvrt_cuda.so", CUDA_SYMBOLS)
    return _cuda


def host_lib():
    global _host
    if _host is None:
        _host = _load("libvrt_host.so", HOST_SYMBOLS)
    return _host
