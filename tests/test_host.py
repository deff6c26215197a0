"""CPU tests of the host-side product code (libvrt_host.so) and of the C-ABI surface of libvrt_cuda.so."""
import os
import re
import struct
import zlib

import numpy as np
import pytest
from oracle_lib import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"\w+)\s*\(", txt)))


def test_cuda_library_exports_every_declared_symbol(pkg):
    lib = pkg._ffi.cuda_lib()  # loads without a GPU; binds every entry of CUDA_SYMBOLS
    declared = _declared("vrt_cuda.h", "vrt_cuda_")
    assert declared, "no declarations found"
    for sym in declared:
        assert hasattr(lib, sym), f"libvrt_cuda.so does not export {sym}"
    assert sorted(pkg._ffi.CUDA_SYMBOLS) == declared
    assert lib.vrt_cuda_abi_version() == 5


def test_host_library_exports_every_declared_symbol(pkg):
    lib = pkg._ffi.host_lib()
    declared = _declared("vrt_host.h", "vrt_host_")
    for sym in declared:
        assert hasattr(lib, sym)
    assert sorted(pkg._ffi.HOST_SYMBOLS) == declared


def test_no_cpu_fallback_without_gpu(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.vrt.VrtCudaError, match="no CUDA device"):
        pkg.vrt.Renderer(0)


def test_struct_layouts_match_header(pkg):
    import ctypes

    assert ctypes.sizeof(pkg._ffi.Frame) == 16 * 4 + 4 * 4 + 8 * 4
    assert ctypes.sizeof(pkg._ffi.Stats) == 3 * 8 + 2 * 4 + 2 * 8 + 4 * 4 + 8 + 8  # ... + terms_saturated + terms_terminated


def test_python_flag_values_match_header(pkg):
    """The ctypes mirror (vrt.py) carries the flag and function-id values of include/vrt_cuda.h; evaluate the header's
    #defines and compare every constant the mirror exposes."""
    txt = open(os.path.join(ROOT, "include", "vrt_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    env = {}
    for name, expr in re.findall(r"^#define (VRT_CUDA_\w+) (.+)$", txt, flags=re.M):
        expr = re.sub(r"(\d+)u\b", r"\1", expr.strip())
        env[name] = eval(expr, {}, env)  # noqa: S307 -- the header is ours; later defines refer to earlier ones
    V = pkg.vrt
    names = [n for n in dir(V) if n.isupper() and not n.startswith("_")]
    checked = 0
    for n in names:
        if "VRT_CUDA_" + n in env:
            assert getattr(V, n) == env["VRT_CUDA_" + n], n
            checked += 1
    assert checked >= 30, checked
    for must in ("MODE1", "MODE4", "MODE5", "MODE8", "NO_SKIP", "DEPTH_WINDOW", "APPROX_ERF_TAYLOR", "APPROX_EXP_SPLINE", "FN_SPLINE_EXP", "LIST_BOUND"):
        assert must in names and "VRT_CUDA_" + must in env, must


def test_grid_scene_is_main_cpp_grid(pkg):
    g = pkg.scenes.grid(4)
    assert g.shape == (16, 10)
    assert np.allclose(g[:, 8], 0.125) and np.allclose(g[:, 9], 1.0) and np.allclose(g[:, 6], 1.0)
    assert np.allclose(g[0, :4], [1, 0, 0, 1]) and np.allclose(g[5, :4], [1 - 5 / 16, 0, 5 / 16, 1])
    assert np.allclose(g[1, 4:6], [-0.75, -0.25])  # i is the x index, j the y index (main.cpp:201)
    with pytest.raises(ValueError):
        pkg.scenes.grid(256)  # u8 loop indices
    ie = pkg.scenes.img_error_grid()
    assert ie.shape == (256, 10) and np.allclose(ie[:, 8], 0.25) and np.allclose(ie[:, 9], 3.0)


def test_camera_matches_oracle(pkg):
    for off, focal, rot in ((-4.0, 1.0, 0.0), (-5.0, 1.3, 77.0), (-2.5, 0.9, 200.0)):
        cam, origin = pkg.vrt.camera_t.app(64, 64, off, focal, rot)
        view_o, origin_o = Oracle.app_camera(off, focal, rot)
        assert np.abs(cam.view_matrix - view_o).max() <= 2e-6
        assert np.abs(origin - origin_o).max() <= 2e-6
    gold = np.load(os.path.join(GOLDEN, "reference_outputs.npz"))
    for row in gold["camera_views"]:
        cam = pkg.vrt.camera_t(row[0:3], float(row[3]), float(row[4]), 32, 32, float(row[5]))
        assert np.abs(cam.view_matrix - row[6:22]).max() <= 2e-6
    row = gold["app_cameras"][0]
    # default camera: proj = (x, y, z + 3)  (SURVEY.md hard part B)
    assert np.abs(row[3:19].reshape(4, 4).T - np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 3], [0, 0, 0, 1]])).max() < 1e-6


def test_read_obj_matches_reference_output(pkg, tmp_path):
    for name in ("sphere", "simple_cube", "monkey"):
        src = np.load(os.path.join(GOLDEN, f"{name}_gaussians.npy"))
        p = tmp_path / f"{name}.obj"
        with open(p, "w") as f:
            f.write("# Blender\nmtllib x.mtl\no Obj\n")
            for g in src:
                f.write("v %.6f %.6f %.6f\n" % (g[4], g[5], g[6]))
            f.write("vn 0.0 1.0 0.0\nvt 0.5 0.5\ns 0\nf 1/1/1 2/1/1 3/1/1\n")
        got = pkg.scenes.read_obj(str(p))
        assert got.shape == src.shape
        assert np.abs(got - src).max() <= 1e-6, name
    sig = {n: float(np.load(os.path.join(GOLDEN, f"{n}_gaussians.npy"))[0, 8]) for n in ("sphere", "cube", "monkey", "teapot")}
    assert sig == {"sphere": pytest.approx(0.3), "cube": pytest.approx(0.15), "monkey": pytest.approx(0.15), "teapot": pytest.approx(0.05)}
    with pytest.raises(OSError):
        pkg.scenes.read_obj(str(tmp_path / "missing.obj"))


def test_synthetic_scene_is_deterministic_and_in_range(pkg):
    a = pkg.scenes.synthetic(5000, 42, -2.1, -1.5)
    b = pkg.scenes.synthetic(5000, 42, -2.1, -1.5)
    c = pkg.scenes.synthetic(5000, 43, -2.1, -1.5)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.array_equal(a[:100], pkg.scenes.synthetic(100, 42, -2.1, -1.5))  # counter based: prefix stable
    z = a[:, 6]
    assert z.min() >= 0 and z.max() <= 2
    assert np.all(np.abs(a[:, 4]) <= z + 4 + 1e-5) and np.all(np.abs(a[:, 5]) <= z + 4 + 1e-5)
    assert a[:, 8].min() >= 10 ** -2.1 * 0.999 and a[:, 8].max() <= 10 ** -1.5 * 1.001
    tau = a[:, 9] * a[:, 8] * np.sqrt(2 * np.pi)
    assert tau.min() >= 0.2 - 1e-4 and tau.max() <= 1.5 + 1e-4
    assert np.all(a[:, 3] == 1) and np.all(a[:, 7] == 0)
    assert a[:, :3].min() >= 0 and a[:, :3].max() < 1


def test_row_bands(pkg):
    import ctypes

    lib = pkg._ffi.host_lib()

    def bands(cost, parts):
        c = np.asarray(cost, np.float64)
        out = np.zeros(parts + 1, np.uint32)
        assert lib.vrt_host_row_bands(c.ctypes.data_as(ctypes.c_void_p), len(c), parts, out.ctypes.data_as(ctypes.c_void_p)) == 0
        return out

    b = bands([1] * 8, 4)
    assert list(b) == [0, 2, 4, 6, 8]
    b = bands([10, 1, 1, 1, 1, 1, 1, 10], 3)
    assert b[0] == 0 and b[-1] == 8 and np.all(np.diff(b.astype(int)) >= 0)
    cost = np.array([10, 1, 1, 1, 1, 1, 1, 10], float)
    worst = max(cost[b[i] : b[i + 1]].sum() for i in range(3))
    assert worst == 10  # optimal: [10] [1 x 6] [10]
    rng = np.random.default_rng(0)
    cost = rng.random(64) ** 4
    b = bands(cost, 8)
    worst = max(cost[b[i] : b[i + 1]].sum() for i in range(8))
    assert worst <= cost.sum() / 8 + cost.max() + 1e-12
    b = bands([0, 0, 5], 2)  # more parts than useful rows still covers every row once
    assert b[0] == 0 and b[-1] == 3
    assert lib.vrt_host_row_bands(None, 3, 2, None) == -1


def test_write_png_roundtrip(pkg, tmp_path):
    import ctypes

    lib = pkg._ffi.host_lib()
    w, h = 37, 11
    rng = np.random.default_rng(3)
    img = rng.integers(0, 2**32, size=(h, w), dtype=np.uint32)
    p = str(tmp_path / "o.png").encode()
    assert lib.vrt_host_write_png(p, w, h, img.ctypes.data_as(ctypes.c_void_p)) == 0
    data = open(p, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, ihdr = 8, b"", None
    while pos < len(data):
        ln, typ = struct.unpack(">I4s", data[pos : pos + 8])
        body = data[pos + 8 : pos + 8 + ln]
        assert struct.unpack(">I", data[pos + 8 + ln : pos + 12 + ln])[0] == zlib.crc32(typ + body)
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        if typ == b"IDAT":
            idat += body
        pos += 12 + ln
    assert ihdr == (w, h, 8, 6, 0, 0, 0)
    raw = zlib.decompress(idat)
    rows = np.frombuffer(raw, np.uint8).reshape(h, 1 + 4 * w)
    assert np.all(rows[:, 0] == 0)
    # bytes are the little-endian u32s: PNG red = B channel of 0xAARRGGBB (main.cpp:306)
    assert np.array_equal(rows[:, 1:].reshape(h, w, 4), img.view(np.uint8).reshape(h, w, 4))
