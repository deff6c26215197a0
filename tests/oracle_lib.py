"""ctypes bindings for the parity checkers under oracle/ -- TEST INFRASTRUCTURE ONLY.

`Oracle`  : oracle/libvrt_oracle.so, our plain-C restatement (oracle/vrt_oracle.c).
`Ref`     : oracle/_ref/libvrt_ref_v{3,4}.so, the UNMODIFIED reference compiled in place
            (oracle/Makefile, oracle/ref_driver.cpp).  Optional: `Ref.available()`.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_f, c_d, c_i, c_u64, vp = ctypes.c_float, ctypes.c_double, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p

# approximation tables (orc_approx_table / ref_approx_table) and the variant code of the radiance entry points
APPROX_FNS = ("spline_erf", "spline_erf_mirror", "taylor_erf", "abramowitz_stegun_erf", "erff", "expf", "fast_exp", "spline_exp")
ERF_IDS = {"exact": 0, "as": 1, "spline": 2, "spline_mirror": 3, "taylor": 4}
EXP_IDS = {"exact": 0, "fast": 1, "spline": 2}


def variant_code(erf="exact", exp="exact"):
    return ERF_IDS[erf] | (EXP_IDS[exp] << 4)



def _ptr(a):
    return a.ctypes.data_as(vp)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


class Oracle:
    """Plain-C restatement; every method cites the reference lines in oracle/vrt_oracle.c."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = os.path.join(ORACLE_DIR, "libvrt_oracle.so")
            if not os.path.exists(path):
                raise RuntimeError(f"{path} missing: run `make -C oracle oracle` (or __graft_entry__.build())")
            L = ctypes.CDLL(path)
            L.orc_as_erf_f32.restype = c_f
            L.orc_as_erf_f32.argtypes = [c_f]
            L.orc_transmittance_f32.restype = c_f
            L.orc_transmittance_f32.argtypes = [vp, vp, c_f, vp, c_u64, c_i]
            L.orc_transmittance_f64.restype = c_d
            L.orc_transmittance_f64.argtypes = [vp, vp, c_d, vp, c_u64, c_i]
            L.orc_radiance_rays_f32.argtypes = [vp, vp, c_u64, vp, c_u64, c_i, vp]
            L.orc_radiance_rays_f64.argtypes = [vp, vp, c_u64, vp, c_u64, c_i, vp]
            L.orc_radiance_rays_f64_unit.argtypes = [vp, vp, c_u64, vp, c_u64, c_i, vp]
            L.orc_view_matrix.argtypes = [vp, c_f, c_f, c_f, vp]
            L.orc_app_camera.argtypes = [c_f, c_f, c_f, vp, vp]
            L.orc_inverse4.argtypes = [vp, vp]
            L.orc_pixel_dirs.argtypes = [vp, vp, c_u64, c_u64, vp, c_u64, vp]
            L.orc_tile_membership.restype = c_u64
            L.orc_tile_membership.argtypes = [c_f, c_f, vp, c_u64, vp, vp, vp, vp, vp, c_u64, vp, c_u64]
            L.orc_pack_pixel.restype = ctypes.c_uint32
            L.orc_pack_pixel.argtypes = [vp, c_i, c_i]
            L.orc_read_obj.restype = c_u64
            L.orc_read_obj.argtypes = [ctypes.c_char_p, vp, c_u64]
            L.orc_approx_table.argtypes = [c_i, vp, c_u64, vp]
            cls._lib = L
        return cls._lib

    @classmethod
    def as_erf(cls, x):
        L = cls.lib()
        return np.array([L.orc_as_erf_f32(float(v)) for v in np.asarray(x, np.float32)], np.float32)

    @classmethod
    def approx_table(cls, fn, x):
        """APPROX_FNS[fn] evaluated at x (the functions tests/accuracy.cpp tabulates)."""
        xv = _f32(x)
        out = np.zeros_like(xv)
        cls.lib().orc_approx_table(fn, _ptr(xv), len(xv), _ptr(out))
        return out

    @classmethod
    def transmittance(cls, gaussians, origin, direction, s, variant=0, f64=False):
        L = cls.lib()
        g, o, d = _f32(gaussians), _f32(origin), _f32(direction)
        fn = L.orc_transmittance_f64 if f64 else L.orc_transmittance_f32
        return np.array([fn(_ptr(o), _ptr(d), float(v), _ptr(g), len(g), variant) for v in s], np.float64 if f64 else np.float32)

    @classmethod
    def radiance(cls, gaussians, origin, dirs, variant=0, f64=False):
        """f64=False: fp32 IEEE evaluation; True: double on the fp32 directions as given; "unit": double on the
        directions re-normalised in double (the arbiter for ill-conditioned scenes)."""
        L = cls.lib()
        g, o, d = _f32(gaussians), _f32(origin), _f32(dirs)
        out = np.zeros((len(d), 4), np.float64 if f64 else np.float32)
        fn = L.orc_radiance_rays_f64_unit if f64 == "unit" else (L.orc_radiance_rays_f64 if f64 else L.orc_radiance_rays_f32)
        fn(_ptr(o), _ptr(d), len(d), _ptr(g), len(g), variant, _ptr(out))
        return out

    @classmethod
    def view_matrix(cls, pos, yaw, pitch, focal):
        L = cls.lib()
        p, out = _f32(pos), np.zeros(16, np.float32)
        L.orc_view_matrix(_ptr(p), yaw, pitch, focal, _ptr(out))
        return out

    @classmethod
    def app_camera(cls, camera_offset=-4.0, focal=1.0, initial_rot=0.0):
        L = cls.lib()
        view, origin = np.zeros(16, np.float32), np.zeros(4, np.float32)
        L.orc_app_camera(camera_offset, focal, initial_rot, _ptr(view), _ptr(origin))
        return view, origin

    @classmethod
    def inverse4(cls, m):
        L = cls.lib()
        a, out = _f32(m).reshape(16), np.zeros(16, np.float32)
        L.orc_inverse4(_ptr(a), _ptr(out))
        return out

    @classmethod
    def pixel_dirs(cls, view, origin, w, h, pix):
        L = cls.lib()
        v, o, p = _f32(view), _f32(origin), np.ascontiguousarray(pix, np.uint64)
        out = np.zeros((len(p), 4), np.float32)
        L.orc_pixel_dirs(_ptr(v), _ptr(o), w, h, _ptr(p), len(p), _ptr(out))
        return out

    @classmethod
    def tile_membership(cls, tw, th, gaussians, view):
        return _membership(cls.lib().orc_tile_membership, tw, th, gaussians, view)

    @classmethod
    def pack_pixel(cls, rgba, round_nearest, alpha_quirk):
        c = _f32(rgba)
        return int(cls.lib().orc_pack_pixel(_ptr(c), int(round_nearest), int(alpha_quirk)))

    @classmethod
    def read_obj(cls, path):
        L = cls.lib()
        n = L.orc_read_obj(path.encode(), None, 0)
        if n == 2**64 - 1:
            raise OSError(path)
        out = np.zeros((n, 10), np.float32)
        L.orc_read_obj(path.encode(), _ptr(out), n)
        return out


def _membership(fn, tw, th, gaussians, view):
    """-> (tiles_w, tiles_h, counts[n_lists], idx[total]); n_lists is the number of tile centres the
    reference's float-accumulation loops produce (normally tiles_w*tiles_h)."""
    g, v = _f32(gaussians), _f32(view)
    tw_, th_, nl_ = c_u64(), c_u64(), c_u64()
    cap_c = 1 << 18
    counts = np.zeros(cap_c, np.uint32)
    args = (np.float32(tw), np.float32(th), _ptr(g), len(g), _ptr(v), ctypes.addressof(tw_), ctypes.addressof(th_), ctypes.addressof(nl_), _ptr(counts), cap_c)
    total = fn(*args, None, 0)
    if total == 2**64 - 1:
        raise RuntimeError("too many tiles")
    idx = np.zeros(max(int(total), 1), np.uint32)
    fn(*args, _ptr(idx), int(total))
    return int(tw_.value), int(th_.value), counts[: int(nl_.value)].copy(), idx[: int(total)].copy()


class Ref:
    """The unmodified reference, compiled in place.  Present in the CPU container (built from
    /root/reference by oracle/Makefile) and shipped to the GPU box as a prebuilt .so."""

    _lib = None
    _tried = False

    @classmethod
    def path(cls):
        flags = _cpu_flags()
        want = []
        if {"avx512f", "avx512bw", "avx512dq", "avx512vl", "avx512cd"} <= flags:
            want.append("libvrt_ref_v4.so")
        if {"avx2", "fma", "bmi2"} <= flags:
            want.append("libvrt_ref_v3.so")
        for name in want:
            p = os.path.join(ORACLE_DIR, "_ref", name)
            if os.path.exists(p):
                return p
        return None

    @classmethod
    def available(cls):
        return cls.lib(required=False) is not None

    @classmethod
    def lib(cls, required=True):
        if cls._lib is None and not cls._tried:
            cls._tried = True
            p = cls.path()
            if p is not None:
                L = ctypes.CDLL(p)
                L.ref_simd_floats.restype = c_i
                L.ref_app_camera.argtypes = [c_f, c_f, c_f, c_u64, c_u64, vp, vp]
                L.ref_camera_view.argtypes = [vp, c_f, c_f, c_f, c_u64, c_u64, vp]
                L.ref_pixel_dirs.argtypes = [vp, c_f, c_f, c_f, c_u64, c_u64, vp, vp, c_u64, vp]
                L.ref_tile_membership.restype = c_u64
                L.ref_tile_membership.argtypes = [c_f, c_f, vp, c_u64, vp, vp, vp, vp, vp, c_u64, vp, c_u64]
                L.ref_radiance.argtypes = [vp, c_u64, vp, vp, c_u64, c_i, vp]
                L.ref_transmittance.argtypes = [vp, c_u64, vp, vp, vp, c_u64, c_i, vp]
                L.ref_as_erf.argtypes = [vp, c_u64, vp]
                L.ref_render_app.restype = c_i
                L.ref_render_app.argtypes = [c_i, vp, c_u64, c_u64, c_u64, c_u64, c_u64, c_f, c_f, c_f, vp, vp, vp]
                L.ref_render_tile_strip.restype = c_d
                L.ref_render_tile_strip.argtypes = [vp, vp, c_u64, c_u64, c_u64, vp, vp, vp, vp, c_u64, c_i, vp]
                L.ref_read_obj.restype = c_u64
                L.ref_read_obj.argtypes = [ctypes.c_char_p, vp, c_u64]
                L.ref_approx_table.restype = c_i
                L.ref_approx_table.argtypes = [c_i, vp, c_u64, vp]
                L.ref_img_error_image.restype = c_i
                L.ref_img_error_image.argtypes = [vp, c_u64, c_f, vp, c_u64, c_u64, c_u64, c_i, vp]
                cls._lib = L
                cls._path = p
        if cls._lib is None and required:
            raise RuntimeError("oracle/_ref/libvrt_ref_*.so missing or not runnable on this CPU: run `make -C oracle ref`")
        return cls._lib

    @classmethod
    def simd_floats(cls):
        return cls.lib().ref_simd_floats()

    @classmethod
    def app_camera(cls, camera_offset=-4.0, focal=1.0, initial_rot=0.0, w=16, h=16):
        view, origin = np.zeros(16, np.float32), np.zeros(4, np.float32)
        cls.lib().ref_app_camera(camera_offset, focal, initial_rot, w, h, _ptr(view), _ptr(origin))
        return view, origin

    @classmethod
    def camera_view(cls, pos, yaw, pitch, focal, w=16, h=16):
        p, out = _f32(pos), np.zeros(16, np.float32)
        cls.lib().ref_camera_view(_ptr(p), yaw, pitch, focal, w, h, _ptr(out))
        return out

    @classmethod
    def pixel_dirs(cls, pos, yaw, pitch, focal, w, h, origin, pix):
        p, o, px = _f32(pos), _f32(origin), np.ascontiguousarray(pix, np.uint64)
        out = np.zeros((len(px), 4), np.float32)
        cls.lib().ref_pixel_dirs(_ptr(p), yaw, pitch, focal, w, h, _ptr(o), _ptr(px), len(px), _ptr(out))
        return out

    @classmethod
    def tile_membership(cls, tw, th, gaussians, view):
        return _membership(cls.lib().ref_tile_membership, tw, th, gaussians, view)

    @classmethod
    def radiance(cls, gaussians, origin, dirs, variant=0):
        g, o, d = _f32(gaussians), _f32(origin), _f32(dirs)
        out = np.zeros((len(d), 4), np.float32)
        cls.lib().ref_radiance(_ptr(g), len(g), _ptr(o), _ptr(d), len(d), variant, _ptr(out))
        return out

    @classmethod
    def transmittance(cls, gaussians, origin, direction, s, variant=0):
        g, o, d, sv = _f32(gaussians), _f32(origin), _f32(direction), _f32(s)
        out = np.zeros(len(sv), np.float32)
        cls.lib().ref_transmittance(_ptr(g), len(g), _ptr(o), _ptr(d), _ptr(sv), len(sv), variant, _ptr(out))
        return out

    @classmethod
    def as_erf(cls, x):
        xv = _f32(x)
        out = np.zeros_like(xv)
        cls.lib().ref_as_erf(_ptr(xv), len(xv), _ptr(out))
        return out

    @classmethod
    def approx_table(cls, fn, x, simd=False):
        xv = _f32(x)
        out = np.zeros_like(xv)
        if cls.lib().ref_approx_table(fn + (16 if simd else 0), _ptr(xv), len(xv), _ptr(out)) != 0:
            raise ValueError(fn)
        return out

    @classmethod
    def img_error_image(cls, gaussians, tw, tiling_view, w, h, variant, threads=4):
        """tests/img-error.cpp:27-43 for one variant code (>= 0: tiled SIMD entry, < 0: the scalar reference image)."""
        g, v = _f32(gaussians), _f32(tiling_view)
        img = np.zeros(w * h, np.uint32)
        cls.lib().ref_img_error_image(_ptr(g), len(g), np.float32(tw), _ptr(v), w, h, threads, variant, _ptr(img))
        return img.reshape(h, w)

    @classmethod
    def render_app(cls, mode, gaussians, w, h, tiles=16, threads=1, camera_offset=-4.0, focal=1.0, initial_rot=0.0):
        """One frame of the reference app (main.cpp:257-297).  Returns image(u32 h*w), (tiling_ms, draw_ms), terms."""
        g = _f32(gaussians)
        img, times, terms = np.zeros(w * h, np.uint32), np.zeros(2), np.zeros(1)
        rc = cls.lib().ref_render_app(mode, _ptr(g), len(g), w, h, tiles, threads, camera_offset, focal, initial_rot, _ptr(img), _ptr(times), _ptr(terms))
        if rc != 0:
            raise RuntimeError(f"ref_render_app rc={rc}")
        return img.reshape(h, w), (float(times[0]), float(times[1])), float(terms[0])

    @classmethod
    def render_tile_strip(cls, lists, tile_w, tile_h, plane, origin, threads, scalar_variant=-1):
        """Time the reference render entry on explicit tiles.  lists: list of (n_i,10) arrays (len must be a
        power of two); plane: (3, tile_h, n_tiles*tile_w) projection-plane points.  Returns (ms, image)."""
        n_tiles = len(lists)
        assert n_tiles & (n_tiles - 1) == 0, "n_tiles must be a power of two (2/(2/n) must be exact)"
        offs = np.zeros(n_tiles + 1, np.uint64)
        offs[1:] = np.cumsum([len(l) for l in lists])
        cat = _f32(np.concatenate([np.asarray(l, np.float32).reshape(-1, 10) for l in lists], 0)) if offs[-1] else np.zeros((1, 10), np.float32)
        w = n_tiles * tile_w
        xs, ys, zs = (_f32(plane[i]).reshape(-1) for i in range(3))
        assert xs.size == w * tile_h
        o = _f32(origin)
        img = np.zeros(w * tile_h, np.uint32)
        ms = cls.lib().ref_render_tile_strip(_ptr(cat), _ptr(offs), n_tiles, tile_w, tile_h, _ptr(xs), _ptr(ys), _ptr(zs), _ptr(o), threads, scalar_variant, _ptr(img))
        return float(ms), img.reshape(tile_h, w)

    @classmethod
    def read_obj(cls, path):
        n = cls.lib().ref_read_obj(path.encode(), None, 0)
        out = np.zeros((n, 10), np.float32)
        cls.lib().ref_read_obj(path.encode(), _ptr(out), n)
        return out
