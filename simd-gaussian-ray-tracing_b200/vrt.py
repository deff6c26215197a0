"""Host-side mirror of the reference's `vrt` interface for the hot path, on top of the C ABI.

Names follow the reference (src/vrt/rt.h, src/vrt/rt.cpp, src/vrt/camera.h, src/volumetric-ray-tracer/main.cpp):

    camera_t(position, yaw, pitch, width, height, focal_length)          camera.h:20-44
    tile_gaussians(tw, th, gaussians, view)            -> tiles_t         rt.cpp:29-69 / rt.h:140
    render_image(w, h, cam, origin, gaussians | tiles) -> image           rt.h:227-310  (scalar: exact erf, truncation)
    simd_render_image(w, h, cam, origin, gaussians | tiles) -> image      rt.h:315-404  (A&S erf, round-to-nearest,
                                                                                          alpha quirk in the tiled form)

A scene is a float32 array of shape (n, 10): the memory layout of gaussian_t (types.h:195-200).
Everything renders on the GPU through libvrt_cuda.so; there is no CPU path in this module.
"""
import ctypes

import numpy as np

from . import _ffi
from ._ffi import Frame, Stats

# flag values of include/vrt_cuda.h
ERF_AS, ERF_EXACT = 0, 1
LIST_REFERENCE, LIST_REFERENCE_BOUND, LIST_ALL, LIST_BOUND = 0 << 2, 1 << 2, 2 << 2, 3 << 2
QUANT_TRUNCATE, QUANT_NEAREST = 0 << 4, 1 << 4
ALPHA_OPAQUE, ALPHA_FROM_W = 0 << 5, 1 << 5
NO_SKIP = 1 << 6
DEPTH_WINDOW = 1 << 7  # round-1 name of the (now default) banded evaluation
EVAL_ALL = 1 << 12  # evaluate every listed term: no saturation shortcut, no early exit
NO_TERMINATE = 1 << 13  # banded evaluation without the transmittance early exit
INTERRUPTED = 1  # vrt_cuda_render_interruptible: `running` went false mid-frame
# alternative approximations (src/vrt/approx.h:10-46) and the function ids of Renderer.approx_table
APPROX_ERF_SPLINE, APPROX_ERF_SPLINE_MIRROR, APPROX_ERF_TAYLOR, APPROX_ERF_MASK = 1 << 8, 2 << 8, 3 << 8, 3 << 8
APPROX_EXP_FAST, APPROX_EXP_SPLINE, APPROX_EXP_MASK = 1 << 10, 2 << 10, 3 << 10
FN_SPLINE_ERF, FN_SPLINE_ERF_MIRROR, FN_TAYLOR_ERF, FN_AS_ERF, FN_ERF, FN_EXP, FN_FAST_EXP, FN_SPLINE_EXP = range(8)
MODE1 = ERF_EXACT | LIST_ALL | QUANT_TRUNCATE | ALPHA_OPAQUE
MODE4 = ERF_AS | LIST_ALL | QUANT_NEAREST | ALPHA_OPAQUE
MODE5 = ERF_EXACT | LIST_REFERENCE | QUANT_TRUNCATE | ALPHA_OPAQUE
MODE8 = ERF_AS | LIST_REFERENCE | QUANT_NEAREST | ALPHA_FROM_W
LIST_MASK = 3 << 2


class VrtCudaError(RuntimeError):
    pass


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class camera_t:
    """camera_t of src/vrt/camera.h:20-44 without the per-pixel projection-plane arrays: the GPU regenerates
    the plane point of a pixel from inverse(view) (camera.cpp:60-70) instead of reading 12 B/pixel."""

    def __init__(self, position=(0.0, 0.0, 0.0), yaw=-90.0, pitch=0.0, width=256, height=256, focal_length=1.0):
        self.position = _f32(position)
        self.yaw, self.pitch = float(yaw), float(pitch)
        self.w, self.h = int(width), int(height)
        self.focal_length = float(focal_length)
        self.update()

    def turn(self, yaw, pitch):
        self.yaw, self.pitch = float(yaw), float(pitch)
        self.update()

    def update(self):
        out = np.zeros(16, np.float32)
        _ffi.host_lib().vrt_host_view_matrix(_ptr(self.position), self.yaw, self.pitch, self.focal_length, _ptr(out))
        self.view_matrix = out

    @classmethod
    def app(cls, width, height, camera_offset=-4.0, focal_length=1.0, rotation=0.0):
        """The orbiting camera of main.cpp:248-255 -> (camera, origin)."""
        view, origin = np.zeros(16, np.float32), np.zeros(4, np.float32)
        _ffi.host_lib().vrt_host_app_camera(camera_offset, focal_length, rotation, _ptr(view), _ptr(origin))
        cam = cls(origin[:3], -90.0 - rotation, 0.0, width, height, focal_length)
        cam.view_matrix = view
        return cam, origin


class tiles_t:
    """Per-tile Gaussian lists as returned by tile_gaussians (types.h:272-287): `counts[t]`, `indices` (concatenated,
    row-major tiles, y outer), plus tw/th/w/h.  `gaussians(t)` materialises tile t's records like tiles_t::gaussians[t]."""

    def __init__(self, scene, counts, indices, tw, th, w, h):
        self.scene, self.counts, self.indices = scene, counts, indices
        self.offsets = np.concatenate([[0], np.cumsum(counts, dtype=np.uint64)]).astype(np.uint64)
        self.tw, self.th, self.w, self.h = tw, th, w, h

    def list(self, t):
        return self.indices[int(self.offsets[t]) : int(self.offsets[t + 1])]

    def gaussians(self, t):
        return self.scene[self.list(t)]


class Renderer:
    """One GPU context (vrt_cuda_ctx): owns the device copy of the scene and the per-frame lists."""

    def __init__(self, device=0):
        self._lib = _ffi.cuda_lib()
        h = ctypes.c_void_p()
        rc = self._lib.vrt_cuda_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise VrtCudaError(f"vrt_cuda_create({device}) = {rc}: {self._lib.vrt_cuda_last_error(None).decode()}")
        self._h = h
        self.device = int(device)
        self.n = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vrt_cuda_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise VrtCudaError(f"{what} = {rc}: {self._lib.vrt_cuda_last_error(self._h).decode()}")

    # ---- scene ----
    def set_gaussians(self, scene):
        g = _f32(scene).reshape(-1, 10)
        self._check(self._lib.vrt_cuda_set_gaussians(self._h, _ptr(g), len(g)), "vrt_cuda_set_gaussians")
        self.n = len(g)
        self._scene = g

    def set_gaussians_device(self, dev_ptr, n):
        self._check(self._lib.vrt_cuda_set_gaussians_device(self._h, ctypes.c_void_p(int(dev_ptr)), int(n)), "vrt_cuda_set_gaussians_device")
        self.n = int(n)

    def set_tuning(self, emitter_block, packed):
        self._check(self._lib.vrt_cuda_set_tuning(self._h, int(emitter_block), int(packed)), "vrt_cuda_set_tuning")

    def set_band_tuning(self, ctas_per_sm):
        self._check(self._lib.vrt_cuda_set_band_tuning(self._h, int(ctas_per_sm)), "vrt_cuda_set_band_tuning")

    # ---- frame ----
    @staticmethod
    def frame(view, origin, width, height, flags, tiles=(1, 1), bound_sigmas=0.0, rows=(0, 0)):
        f = Frame()
        v, o = _f32(view).reshape(16), _f32(origin).reshape(-1)
        for i in range(16):
            f.view[i] = float(v[i])
        for i in range(4):
            f.origin[i] = float(o[i]) if i < len(o) else 0.0
        f.width, f.height = int(width), int(height)
        f.tiles_x, f.tiles_y = int(tiles[0]), int(tiles[1])
        f.flags = int(flags)
        f.bound_sigmas = float(bound_sigmas)
        f.row_begin, f.row_end = int(rows[0]), int(rows[1])
        return f

    def tile(self, frame):
        self._check(self._lib.vrt_cuda_tile(self._h, ctypes.byref(frame)), "vrt_cuda_tile")

    def set_tile_lists(self, frame, lists):
        """lists: sequence of (n_t, 10) arrays, one per reference tile (row-major, y outer) -- a tiles_t."""
        offs = np.zeros(len(lists) + 1, np.uint64)
        offs[1:] = np.cumsum([len(l) for l in lists])
        cat = _f32(np.concatenate([np.asarray(l, np.float32).reshape(-1, 10) for l in lists], 0)) if offs[-1] else np.zeros((1, 10), np.float32)
        self._check(self._lib.vrt_cuda_set_tile_lists(self._h, ctypes.byref(frame), _ptr(cat), _ptr(offs), len(lists)), "vrt_cuda_set_tile_lists")

    def get_lists(self):
        nc, ne = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self._lib.vrt_cuda_get_lists(self._h, None, 0, None, 0, ctypes.byref(nc), ctypes.byref(ne)), "vrt_cuda_get_lists")
        counts, idx = np.zeros(max(nc.value, 1), np.uint32), np.zeros(max(ne.value, 1), np.uint32)
        self._check(self._lib.vrt_cuda_get_lists(self._h, _ptr(counts), len(counts), _ptr(idx), len(idx), None, None), "vrt_cuda_get_lists")
        return counts[: nc.value], idx[: ne.value]

    def row_costs(self):
        n, hpx = ctypes.c_uint32(), ctypes.c_uint32()
        self._check(self._lib.vrt_cuda_row_costs(self._h, None, 0, ctypes.byref(n), ctypes.byref(hpx)), "vrt_cuda_row_costs")
        rows = np.zeros(n.value, np.float64)
        self._check(self._lib.vrt_cuda_row_costs(self._h, _ptr(rows), n.value, None, None), "vrt_cuda_row_costs")
        return rows, int(hpx.value)

    def render(self, frame, want_image=True, want_radiance=False, image=None, radiance=None):
        """Render with the current lists into host arrays -> (image u32 [h,w] | None, radiance f32 [h,w,4] | None, stats)."""
        h, w = frame.height, frame.width
        if want_image and image is None:
            image = np.zeros((h, w), np.uint32)
        if want_radiance and radiance is None:
            radiance = np.zeros((h, w, 4), np.float32)
        st = Stats()
        self._check(self._lib.vrt_cuda_render(self._h, ctypes.byref(frame), _ptr(image) if image is not None else None,
                                              _ptr(radiance) if radiance is not None else None, ctypes.byref(st)), "vrt_cuda_render")
        return image, radiance, st.as_dict()

    def render_interruptible(self, frame, running, want_image=True, want_radiance=False):
        """vrt_cuda_render_interruptible: `running` is a one-byte numpy array (the reference's `const bool &running`) another
        thread may clear.  -> (interrupted: bool, image | None, radiance | None, stats)."""
        h, w = frame.height, frame.width
        image = np.zeros((h, w), np.uint32) if want_image else None
        radiance = np.zeros((h, w, 4), np.float32) if want_radiance else None
        st = Stats()
        rc = self._lib.vrt_cuda_render_interruptible(self._h, ctypes.byref(frame), _ptr(image) if image is not None else None,
                                                     _ptr(radiance) if radiance is not None else None, ctypes.byref(st), _ptr(running))
        if rc not in (0, INTERRUPTED):
            self._check(rc, "vrt_cuda_render_interruptible")
        return rc == INTERRUPTED, image, radiance, st.as_dict()

    def abort(self, on=True):
        self._check(self._lib.vrt_cuda_abort(self._h, int(bool(on))), "vrt_cuda_abort")

    def set_host_pinning(self, on=True):
        self._check(self._lib.vrt_cuda_set_host_pinning(self._h, int(bool(on))), "vrt_cuda_set_host_pinning")

    def pin_buffer(self, array):
        self._check(self._lib.vrt_cuda_pin_buffer(self._h, _ptr(array), array.nbytes), "vrt_cuda_pin_buffer")

    def unpin_buffer(self, array):
        self._check(self._lib.vrt_cuda_unpin_buffer(self._h, _ptr(array)), "vrt_cuda_unpin_buffer")

    # multi-GPU output without a gather: a frame buffer on one GPU that other processes' render kernels store into
    def peer_image_create(self, nbytes):
        """(device pointer, 64-byte handle) of a new image on this context's GPU (vrt_cuda_peer_image_create)."""
        ptr, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        self._check(self._lib.vrt_cuda_peer_image_create(self._h, int(nbytes), ctypes.byref(ptr), ctypes.cast(handle, ctypes.c_void_p)), "vrt_cuda_peer_image_create")
        return int(ptr.value), bytes(handle)

    def peer_image_open(self, handle):
        """device pointer, valid on this context's GPU, of an image another process created (vrt_cuda_peer_image_open)."""
        ptr, buf = ctypes.c_void_p(), (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        self._check(self._lib.vrt_cuda_peer_image_open(self._h, ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(ptr)), "vrt_cuda_peer_image_open")
        return int(ptr.value)

    def peer_image_close(self, ptr):
        self._check(self._lib.vrt_cuda_peer_image_close(self._h, ctypes.c_void_p(ptr)), "vrt_cuda_peer_image_close")

    def render_device(self, frame, image_ptr, radiance_ptr=0, want_stats=False):
        st = Stats() if want_stats else None
        self._check(self._lib.vrt_cuda_render_device(self._h, ctypes.byref(frame), ctypes.c_void_p(int(image_ptr)) if image_ptr else None,
                                                     ctypes.c_void_p(int(radiance_ptr)) if radiance_ptr else None,
                                                     ctypes.byref(st) if st is not None else None), "vrt_cuda_render_device")
        return st.as_dict() if st is not None else None

    def frame_render(self, frame, want_image=True, want_radiance=False):
        self.tile(frame)
        return self.render(frame, want_image, want_radiance)

    def approx_table(self, fn, x):
        """y = f(x) on the device for one of the FN_* functions (the columns tests/accuracy.cpp tabulates)."""
        xv = _f32(x).reshape(-1)
        out = np.zeros_like(xv)
        self._check(self._lib.vrt_cuda_approx_table(self._h, int(fn), xv.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), len(xv)),
                    "vrt_cuda_approx_table")
        return out

    def approx_rate(self, fn):
        """values/s of one FN_* device function (the GPU counterpart of tests/approx_cycles.cpp)."""
        out = ctypes.c_double()
        self._check(self._lib.vrt_cuda_approx_rate(self._h, int(fn), ctypes.byref(out)), "vrt_cuda_approx_rate")
        return float(out.value)

    def set_slice(self, slice_):
        self._check(self._lib.vrt_cuda_set_slice(self._h, int(slice_)), "vrt_cuda_set_slice")

    def auto_slice(self, share=1.0):
        out = ctypes.c_int()
        self._check(self._lib.vrt_cuda_auto_slice(self._h, float(share), ctypes.byref(out)), "vrt_cuda_auto_slice")
        return int(out.value)

    def fp32_peak(self, packed=False):
        out = ctypes.c_double()
        self._check(self._lib.vrt_cuda_fp32_peak(self._h, int(packed), ctypes.byref(out)), "vrt_cuda_fp32_peak")
        return float(out.value)

    def term_peak(self, pairs=20, ctas_per_sm=1):
        out = ctypes.c_double()
        self._check(self._lib.vrt_cuda_term_peak(self._h, int(pairs), int(ctas_per_sm), ctypes.byref(out)), "vrt_cuda_term_peak")
        return float(out.value)

    def mix_peak(self, nf, nm, nl):
        out = ctypes.c_double()
        self._check(self._lib.vrt_cuda_mix_peak(self._h, int(nf), int(nm), int(nl), ctypes.byref(out)), "vrt_cuda_mix_peak")
        return float(out.value)

    def sync(self):
        self._check(self._lib.vrt_cuda_sync(self._h), "vrt_cuda_sync")

    @property
    def stream(self):
        return int(self._lib.vrt_cuda_stream(self._h))


_default = {}


def _renderer(device=0):
    if device not in _default:
        _default[device] = Renderer(device)
    return _default[device]


def tile_gaussians(tw, th, gaussians, view, width=256, height=256, device=0):
    """vrt::tile_gaussians(tw, th, gaussians, view) (rt.cpp:29-69): membership of every reference tile, built on the
    GPU by K1.  Returns a tiles_t.  (The image size only fixes the cell grid; membership does not depend on it.)"""
    tx, ty = int(round(2.0 / tw)), int(round(2.0 / th))
    r = _renderer(device)
    scene = _f32(gaussians).reshape(-1, 10)
    r.set_gaussians(scene)
    f = Renderer.frame(view, (0, 0, 0, 0), tx * max(8, width // tx), ty * max(4, height // ty), MODE5, (tx, ty))
    r.tile(f)
    counts, idx = r.get_lists()
    return tiles_t(scene, counts, idx, np.float32(tw), np.float32(th), tx, ty)


def _render(width, height, cam, origin, scene_or_tiles, flags_tiled, flags_untiled, device, want_radiance, list_mode):
    r = _renderer(device)
    if isinstance(scene_or_tiles, tiles_t):
        t = scene_or_tiles
        flags = flags_tiled
        if list_mode is not None:
            flags = (flags & ~LIST_MASK) | list_mode
        f = Renderer.frame(cam.view_matrix, origin, width, height, flags, (t.w, t.h))
        r.set_tile_lists(f, [t.gaussians(i) for i in range(t.w * t.h)])
    else:
        flags = flags_untiled
        if list_mode is not None:
            flags = (flags & ~LIST_MASK) | list_mode
        f = Renderer.frame(cam.view_matrix, origin, width, height, flags)
        r.set_gaussians(scene_or_tiles)
        r.tile(f)
    img, rad, st = r.render(f, True, want_radiance)
    return (img, rad, st) if want_radiance else img


def render_image(width, height, cam, origin, gaussians_or_tiles, device=0, want_radiance=False, list_mode=None):
    """vrt::render_image<radiance<transmittance>> -- the scalar entries (rt.h:227-247 untiled = mode 1, :251-310 tiled =
    mode 5): libm-class exp/erf, truncating quantisation, opaque alpha.  Returns the packed u32 image [h, w]."""
    return _render(width, height, cam, origin, gaussians_or_tiles, MODE5, MODE1, device, want_radiance, list_mode)


def simd_render_image(width, height, cam, origin, gaussians_or_tiles, device=0, want_radiance=False, list_mode=None):
    """vrt::simd_render_image -- the SIMD-over-pixels entries (rt.h:315-337 untiled = mode 4, :344-404 tiled = mode 8):
    Abramowitz-Stegun erf, round-to-nearest quantisation, alpha from colour.w in the tiled form."""
    return _render(width, height, cam, origin, gaussians_or_tiles, MODE8, MODE4, device, want_radiance, list_mode)
