"""Generates tests/golden/*.npy|npz from the UNMODIFIED reference compiled in place (oracle/_ref).

Run in the CPU container (needs /root/reference and `make -C oracle ref`):   python tests/golden/make_golden.py
Only OUTPUTS of the reference are stored (parsed scenes, views, memberships, radiance, images) -- never its sources.
The golden files pin oracle/vrt_oracle.c (tests/test_oracle.py) and feed the GPU tests on boxes without the reference tree.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle_lib import APPROX_FNS, Ref, variant_code  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

REF_OBJ = "/root/reference/test-objects"


def main():
    assert Ref.available(), "build oracle/_ref first"
    pkg = ge.load_package()
    # 1. OBJ scenes as parsed by read_from_obj (gaussians-from-file.cpp:7-44)
    for name in ("sphere", "simple_cube", "cube", "monkey", "teapot"):
        g = Ref.read_obj(os.path.join(REF_OBJ, name + ".obj"))
        np.save(os.path.join(HERE, f"{name}_gaussians.npy"), g)
        print(name, g.shape)

    out = {}
    # 2. cameras: app camera (main.cpp:248-255) at several rotations / focal lengths, and camera_t views
    cams = []
    for off, focal, rot in ((-4.0, 1.0, 0.0), (-4.0, 1.0, 30.0), (-6.0, 1.5, 123.4), (-3.0, 0.8, 270.0)):
        view, origin = Ref.app_camera(off, focal, rot, 16, 16)
        cams.append(np.concatenate([[off, focal, rot], view, origin]))
    out["app_cameras"] = np.array(cams, np.float32)
    views = []
    for pos, yaw, pitch, focal in (((0, 0, 0), -90.0, 0.0, 1.0), ((1, 2, -3), -60.0, 20.0, 1.2), ((0.5, -1, 4), 100.0, -95.0, 0.7)):
        views.append(np.concatenate([pos, [yaw, pitch, focal], Ref.camera_view(pos, yaw, pitch, focal)]))
    out["camera_views"] = np.array(views, np.float32)

    # 3. A&S erf (approx.cpp:90-99) on the grid of tests/accuracy.cpp (x in [-6, 6] step 0.1)
    x = np.arange(-6.0, 6.0001, 0.1, dtype=np.float32)
    out["erf_x"] = x
    out["erf_as"] = Ref.as_erf(x)

    # 4. transmittance sweep of tests/transmittance.cpp:9,24-32
    tg = pkg.scenes.transmittance_test()
    origin = np.array([0, 0, -5, 0], np.float32)
    d = np.array([0, 0, 1, 0], np.float32)
    ks = np.arange(-6.0, 6.0001, 0.1, dtype=np.float32)
    s = ((tg[2, 4:8] - origin) @ d + ks * tg[2, 8]).astype(np.float32)
    out["tr_s"] = s
    out["tr_T_exact"] = Ref.transmittance(tg, origin, d, s, 0)
    out["tr_T_as"] = Ref.transmittance(tg, origin, d, s, 1)

    # 5. config 1 (-g 4, 256x256, 16 tiles): membership, pixel directions, scalar radiance (both erf variants), images
    scene = pkg.scenes.grid(4)
    view, origin = Ref.app_camera(-4.0, 1.0, 0.0, 256, 256)
    tw = np.float32(2.0) / np.float32(16)
    w, h, counts, idx = Ref.tile_membership(tw, tw, scene, view)
    out["c1_counts"], out["c1_idx"] = counts, idx
    pix = np.arange(0, 256 * 256, 53, dtype=np.uint64)
    out["c1_pix"] = pix
    dirs = Ref.pixel_dirs(origin[:3], -90.0, 0.0, 1.0, 256, 256, origin, pix)
    out["c1_dirs"] = dirs
    offs = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)])
    for variant, key in ((0, "c1_rad_exact"), (1, "c1_rad_as")):
        rad = np.zeros((len(pix), 4), np.float32)
        for k, p in enumerate(pix):
            t = (int(p) // 256 // 16) * 16 + (int(p) % 256) // 16
            rad[k] = Ref.radiance(scene[idx[offs[t] : offs[t + 1]]], origin, dirs[k : k + 1], variant)[0]
        out[key] = rad
    out["c1_rad_untiled_as"] = Ref.radiance(scene, origin, dirs, 1)
    for mode in (1, 2, 3, 4, 5, 6, 7, 8):
        img, _, terms = Ref.render_app(mode, scene, 256, 256, tiles=16, threads=8)
        out[f"c1_image_mode{mode}"] = img
        out[f"c1_terms_mode{mode}"] = np.array([terms])

    # 6. img-error.cpp scene: membership with the identity view, tw = 1/8
    g16 = pkg.scenes.img_error_grid()
    ident = np.eye(4, dtype=np.float32).reshape(16)
    w, h, counts, idx = Ref.tile_membership(np.float32(1.0 / 8.0), np.float32(1.0 / 8.0), g16, ident)
    out["ie_counts"], out["ie_idx"] = counts, idx

    # 7. teapot / cube membership (16 tiles, default camera) and a few radiance pixels of the cube
    for name in ("teapot", "cube"):
        g = np.load(os.path.join(HERE, f"{name}_gaussians.npy"))
        w, h, counts, idx = Ref.tile_membership(tw, tw, g, view)
        out[f"{name}_counts"], out[f"{name}_idx"] = counts, idx.astype(np.uint16 if len(g) < 65536 else np.uint32)
    g = np.load(os.path.join(HERE, "cube_gaussians.npy"))
    counts, idx = out["cube_counts"], out["cube_idx"].astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)])
    pix = np.array([r * 256 + c for r in (40, 100, 128, 131, 200) for c in (37, 90, 128, 133, 210)], np.uint64)
    dirs = Ref.pixel_dirs(origin[:3], -90.0, 0.0, 1.0, 256, 256, origin, pix)
    rad = np.zeros((len(pix), 4), np.float32)
    for k, p in enumerate(pix):
        t = (int(p) // 256 // 16) * 16 + (int(p) % 256) // 16
        rad[k] = Ref.radiance(g[idx[offs[t] : offs[t + 1]]], origin, dirs[k : k + 1], 1)[0]
    out["cube_pix"], out["cube_dirs"], out["cube_rad_as"] = pix, dirs, rad

    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_outputs.npz"), {k: v.shape for k, v in out.items()})
    approx_golden(pkg)


# variant combinations (erf, exp) pinned by radiance fixtures; the first two are the FOG / MINE columns of tests/img-error.cpp
APPROX_VARIANTS = (("as", "exact"), ("as", "fast"), ("spline", "exact"), ("spline_mirror", "exact"), ("taylor", "exact"),
                   ("exact", "spline"), ("spline", "fast"), ("taylor", "spline"))


def approx_golden(pkg):
    """8. the alternative approximations (src/vrt/approx.h:10-46): tables on the grids of tests/accuracy.cpp (plus every
    spline knot and its neighbours), scalar-path radiance of the img-error scene for several <Exp, Erf> combinations, and
    the MSE figures tests/img-error.cpp prints (SIMD tiled entry vs the scalar exact image)."""
    out = {}
    knots_e = np.array([-2.9 + 0.6 * i for i in range(11)], np.float32)
    knots_x = np.array([-9, -8, -7, -6, -5, -4.5, -4, -3.5, -3, -2.5, -2, -1.75, -1.5, -1.25, -1, -0.75, -0.5, -0.25, 0], np.float32)
    around = lambda k: np.concatenate([k, np.nextafter(k, np.float32(-100)), np.nextafter(k, np.float32(100))])
    xe = np.concatenate([np.arange(-6.0, 6.0001, 0.1, dtype=np.float32), np.arange(-3.3, 3.3, 0.0137, dtype=np.float32), around(knots_e),
                         np.array([-2.0, 2.0, 0.0, -0.0], np.float32)])
    xx = np.concatenate([np.arange(-16.0, 0.0001, 0.1, dtype=np.float32), np.arange(-10.0, 0.4, 0.0171, dtype=np.float32), around(knots_x),
                         np.array([-87.0, -88.5, -100.0, -1000.0], np.float32)])
    out["erf_x"], out["exp_x"] = xe, xx
    for fn, name in enumerate(APPROX_FNS):
        out[f"table_{name}"] = Ref.approx_table(fn, xe if fn < 5 else xx)
    out["table_fast_exp_simd"] = Ref.approx_table(6, xx, simd=True)

    g16 = pkg.scenes.img_error_grid()
    ident = np.eye(4, dtype=np.float32).reshape(16)
    origin = np.zeros(4, np.float32)
    w, h, counts, idx = Ref.tile_membership(np.float32(1.0 / 8.0), np.float32(1.0 / 8.0), g16, ident)
    offs = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)])
    pix = np.arange(11, 256 * 256, 397, dtype=np.uint64)
    dirs = Ref.pixel_dirs((0, 0, 0), -90.0, 0.0, 1.0, 256, 256, origin, pix)
    out["ie_pix"], out["ie_dirs"] = pix, dirs
    for erf, exp in APPROX_VARIANTS:
        rad = np.zeros((len(pix), 4), np.float32)
        for k, p in enumerate(pix):
            t = (int(p) // 256 // 16) * 16 + (int(p) % 256) // 16
            rad[k] = Ref.radiance(g16[idx[offs[t] : offs[t + 1]]], origin, dirs[k : k + 1], variant_code(erf, exp))[0]
        out[f"ie_rad_{erf}_{exp}"] = rad
        print("radiance", erf, exp, float(rad.max()))
    rgb = lambda im: np.stack([(im >> s) & 0xFF for s in (0, 8, 16)], -1).astype(np.float64) / 255.0
    ref_img = Ref.img_error_image(g16, 1.0 / 8.0, ident, 256, 256, -1, threads=8)
    out["ie_image_scalar"] = ref_img[::5, ::5].copy()
    for erf, exp in APPROX_VARIANTS:
        if erf == "exact":
            continue  # no SIMD exact erf without SVML
        img = Ref.img_error_image(g16, 1.0 / 8.0, ident, 256, 256, variant_code(erf, exp), threads=8)
        out[f"ie_mse_{erf}_{exp}"] = np.array([np.mean(np.sum((rgb(ref_img) - rgb(img)) ** 2, -1))])
        out[f"ie_image_{erf}_{exp}"] = img[::5, ::5].copy()
        print("img-error MSE", erf, exp, float(out[f"ie_mse_{erf}_{exp}"][0]))
    np.savez_compressed(os.path.join(HERE, "reference_approx.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_approx.npz"))


if __name__ == "__main__":
    main()
