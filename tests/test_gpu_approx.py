"""GPU parity tests of the alternative approximations: the reference's spline / Taylor erf and
fast / spline exp (src/vrt/approx.h:10-46) as selectable device functions, called through the C ABI.

Three layers, as in the reference's own tests:
  * tests/accuracy.cpp -- the functions themselves on its x grids (vrt_cuda_approx_table) against the reference's values;
  * the scalar path radiance<transmittance<Exp, Erf>> (rt.h:32, 146) of the img-error scene with each combination;
  * tests/img-error.cpp -- the MSE of each variant's 8-bit image against the scalar exact image.
Tolerance for radiance: max abs <= 1e-3 and PSNR >= 60 dB against the scalar path with the same <Exp, Erf>.
"""
import os

import numpy as np
import pytest
from oracle_lib import APPROX_FNS, Oracle, Ref, variant_code
from parity_util import APPROX_TOL, psnr, reference_lists, table_mismatch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_ABS, TOL_PSNR = 1e-3, 60.0


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "reference_approx.npz"))


def flags_of(V, erf, exp, base):
    f = base & ~1
    f |= {"exact": V.ERF_EXACT, "as": V.ERF_AS, "spline": V.APPROX_ERF_SPLINE, "spline_mirror": V.APPROX_ERF_SPLINE_MIRROR, "taylor": V.APPROX_ERF_TAYLOR}[erf]
    f |= {"exact": 0, "fast": V.APPROX_EXP_FAST, "spline": V.APPROX_EXP_SPLINE}[exp]
    return f


# device tolerance on top of APPROX_TOL: MUFU.EX2 on x log2e loses |x| ulp(x log2e) in the argument
DEVICE_TOL = dict(APPROX_TOL, expf=(1e-37, 8e-6))


def test_tables_match_reference(pkg, renderer, gold):
    """tests/accuracy.cpp's tabulation on the device: every function, the test's grids plus all spline knots and their
    float neighbours (segment selection must be exact: a wrong segment shows as a jump of up to 0.09)."""
    V = pkg.vrt
    ids = (V.FN_SPLINE_ERF, V.FN_SPLINE_ERF_MIRROR, V.FN_TAYLOR_ERF, V.FN_AS_ERF, V.FN_ERF, V.FN_EXP, V.FN_FAST_EXP, V.FN_SPLINE_EXP)
    for fn, name in zip(ids, APPROX_FNS):
        x = gold["erf_x"] if fn < 5 else gold["exp_x"]
        got = renderer.approx_table(fn, x)
        want_ref = gold["table_fast_exp_simd"] if name == "fast_exp" else gold[f"table_{name}"]
        want_orc = Oracle.approx_table(8 if name == "fast_exp" else fn, x)
        a, r = DEVICE_TOL[name]
        for what, want in (("reference", want_ref), ("oracle", want_orc)):
            viol = float((np.abs(got.astype(np.float64) - want) - (a + r * np.abs(want))).max())
            print(f"{name} vs {what}: max abs diff {float(np.abs(got - want).max()):.3e}")
            assert viol <= 0, (name, what, viol)
    assert table_mismatch("erff", renderer.approx_table(V.FN_ERF, gold["erf_x"]), gold["table_erff"], gold["erf_x"]) <= 0


def _ie_setup(pkg, renderer):
    V = pkg.vrt
    scene = pkg.scenes.img_error_grid()
    ident = np.eye(4, dtype=np.float32).reshape(16)
    origin = np.zeros(4, np.float32)
    cam = V.camera_t((0, 0, 0), -90.0, 0.0, 256, 256, 1.0)
    lists = reference_lists(scene, ident, 16)
    renderer.set_gaussians(scene)
    return V, scene, origin, cam, lists


VARIANTS = [("as", "fast"), ("spline", "exact"), ("spline_mirror", "exact"), ("taylor", "exact"), ("exact", "spline"), ("spline", "fast"),
            ("taylor", "spline"), ("as", "spline"), ("exact", "fast"), ("spline_mirror", "fast"), ("spline_mirror", "spline"), ("taylor", "fast"),
            ("spline", "spline")]


@pytest.mark.parametrize("erf,exp", VARIANTS)
def test_variant_radiance(pkg, renderer, gold, erf, exp):
    """img-error scene (tiles from tile_gaussians(1/8, 1/8, grid16, mat4(1)), default camera at the origin) rendered with
    <Exp, Erf> substituted, against the scalar path with the same pair: restatement on every 7th fixture pixel, the
    reference's own values (fixtures generated from oracle/_ref) where a fixture exists."""
    V, scene, origin, cam, lists = _ie_setup(pkg, renderer)
    f = renderer.frame(cam.view_matrix, origin, 256, 256, flags_of(V, erf, exp, V.MODE5), (16, 16))
    renderer.set_tile_lists(f, [scene[l] for l in lists])
    _, rad, st = renderer.render(f, False, True)
    rad = rad.reshape(-1, 4)
    assert np.isfinite(rad).all()
    pix, dirs = gold["ie_pix"].astype(np.int64), gold["ie_dirs"]
    sel = np.arange(0, len(pix), 7)
    want = np.zeros((len(sel), 4), np.float32)
    for i, k in enumerate(sel):
        p = int(pix[k])
        t = (p // 256 // 16) * 16 + (p % 256) // 16
        want[i] = Oracle.radiance(scene[lists[t]], origin, dirs[k : k + 1], variant_code(erf, exp))[0]
    err = float(np.abs(rad[pix[sel]] - want).max())
    print(f"{erf}/{exp} vs restatement: max abs {err:.3e}, PSNR {psnr(rad[pix[sel]], want):.1f} dB, peak {float(want.max()):.3f}, "
          f"executed {st['terms_executed']:.3e} of {st['terms_listed']:.3e} terms")
    assert err <= TOL_ABS and psnr(rad[pix[sel]], want) >= TOL_PSNR
    key = f"ie_rad_{erf}_{exp}"
    if key in gold.files:
        err = float(np.abs(rad[pix] - gold[key]).max())
        print(f"{erf}/{exp} vs reference fixtures: max abs {err:.3e}, PSNR {psnr(rad[pix], gold[key]):.1f} dB")
        assert err <= TOL_ABS and psnr(rad[pix], gold[key]) >= TOL_PSNR


def test_img_error_variant_comparison(pkg, renderer, gold):
    """tests/img-error.cpp:34-58 on the GPU: reference image = tiled scalar exact (mode-5 flags), test images = tiled SIMD
    entry flags (mode 8) with each <Exp, Erf>; metric = mean over pixels of the squared RGB difference of the 8-bit images.
    The GPU's figures must reproduce the reference's (fixtures) -- same ranking, same magnitude."""
    V, scene, origin, cam, lists = _ie_setup(pkg, renderer)
    rgb = lambda im: np.stack([(im >> s) & 0xFF for s in (0, 8, 16)], -1).astype(np.float64) / 255.0

    def image(flags):
        f = renderer.frame(cam.view_matrix, origin, 256, 256, flags, (16, 16))
        renderer.set_tile_lists(f, [scene[l] for l in lists])
        return renderer.render(f, True, False)[0]

    ref_img = image(V.MODE5)
    sub = gold["ie_image_scalar"]
    d = np.abs(rgb(ref_img[::5, ::5]) - rgb(sub)).max() * 255
    assert d <= 1, f"scalar exact image differs from the reference's by {d} LSB"
    got = {}
    for key in [k for k in gold.files if k.startswith("ie_mse_")]:
        erf, exp = key[len("ie_mse_"):].rsplit("_", 1)
        img = image(flags_of(V, erf, exp, V.MODE8))
        mse = float(np.mean(np.sum((rgb(ref_img) - rgb(img)) ** 2, -1)))
        want = float(gold[key][0])
        got[(erf, exp)] = mse
        print(f"img-error MSE {erf}/{exp}: GPU {mse:.4e}, reference {want:.4e}")
        assert abs(mse - want) <= 0.1 * want + 2e-6, (erf, exp, mse, want)
        d = np.abs(rgb(img[::5, ::5]) - rgb(gold[f"ie_image_{erf}_{exp}"])).max() * 255
        assert d <= 2, f"{erf}/{exp}: image differs from the reference's by {d} LSB"
    # the ranking the test exists to show: A&S is two orders of magnitude closer to the exact image than the splines
    assert got[("as", "exact")] < got[("taylor", "exact")] < got[("spline", "exact")]


def test_variants_on_device_built_lists(pkg, renderer):
    """The variant kernel on K1's own index lists (bounded, depth-sorted, split cells): config 1 and the cube object."""
    V = pkg.vrt
    for scene, step in ((pkg.scenes.grid(4), 41), (np.load(os.path.join(GOLDEN, "cube_gaussians.npy")), 331)):
        cam, origin = V.camera_t.app(256, 256)
        renderer.set_gaussians(scene)
        flags = (flags_of(V, "spline_mirror", "fast", V.MODE8) & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND
        f = renderer.frame(cam.view_matrix, origin, 256, 256, flags, (16, 16))
        _, rad, st = renderer.frame_render(f, False, True)
        counts, idx = renderer.get_lists()
        offs = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)])
        pix = np.arange(5, 256 * 256, step, dtype=np.uint64)
        dirs = Oracle.pixel_dirs(cam.view_matrix, origin, 256, 256, pix)
        ncx = 256 // 8
        want = np.zeros((len(pix), 4), np.float32)
        for k, p in enumerate(pix.astype(np.int64)):
            c = (p // 256 // 4) * ncx + (p % 256) // 8
            lst = scene[np.sort(idx[offs[c] : offs[c + 1]])]
            if len(lst):
                want[k] = Oracle.radiance(lst, origin, dirs[k : k + 1], variant_code("spline_mirror", "fast"))[0]
        got = rad.reshape(-1, 4)[pix.astype(np.int64)]
        err = float(np.abs(got - want).max())
        print(f"n={len(scene)}: spline_mirror/fast on bounded lists: max abs {err:.3e}, max list {st['max_list']}")
        assert err <= TOL_ABS and psnr(got, want) >= TOL_PSNR


def test_variant_errors(pkg, renderer):
    V = pkg.vrt
    scene = pkg.scenes.grid(4)
    cam, origin = V.camera_t.app(256, 256)
    renderer.set_gaussians(scene)
    flags = (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND | V.DEPTH_WINDOW | V.APPROX_ERF_SPLINE
    f = renderer.frame(cam.view_matrix, origin, 256, 256, flags, (16, 16))
    renderer.tile(f)
    with pytest.raises(V.VrtCudaError):
        renderer.render(f, True, False)
    f = renderer.frame(cam.view_matrix, origin, 256, 256, V.MODE8 | (3 << 10), (16, 16))
    renderer.tile(f)
    with pytest.raises(V.VrtCudaError):
        renderer.render(f, True, False)
    with pytest.raises(V.VrtCudaError):
        renderer.approx_table(99, np.zeros(4, np.float32))
    assert len(renderer.approx_table(V.FN_AS_ERF, np.zeros(0, np.float32))) == 0


@pytest.mark.skipif(not Ref.available(), reason="compiled reference (oracle/_ref) not present")
def test_tables_against_compiled_reference_live(pkg, renderer):
    """Random arguments (not just the fixture grids) against the reference's own functions, run here."""
    V = pkg.vrt
    rng = np.random.default_rng(7)
    xe = rng.uniform(-4, 4, 20000).astype(np.float32)
    xx = rng.uniform(-30, 0.5, 20000).astype(np.float32)
    ids = (V.FN_SPLINE_ERF, V.FN_SPLINE_ERF_MIRROR, V.FN_TAYLOR_ERF, V.FN_AS_ERF, V.FN_ERF, V.FN_EXP, V.FN_FAST_EXP, V.FN_SPLINE_EXP)
    for fn, name in zip(ids, APPROX_FNS):
        x = xe if fn < 5 else xx
        got = renderer.approx_table(fn, x)
        want = Ref.approx_table(fn, x, simd=(name == "fast_exp"))
        a, r = DEVICE_TOL[name]
        viol = float((np.abs(got.astype(np.float64) - want) - (a + r * np.abs(want))).max())
        assert viol <= 0, (name, viol)


def test_function_throughput_probe(pkg, renderer):
    """tests/approx_cycles.cpp counts CPU cycles per value of every approximation; vrt_cuda_approx_rate is its GPU
    counterpart (values/s, register-resident argument chains).  Only sanity is asserted; the figures go to profiles/."""
    V = pkg.vrt
    ids = (V.FN_SPLINE_ERF, V.FN_SPLINE_ERF_MIRROR, V.FN_TAYLOR_ERF, V.FN_AS_ERF, V.FN_ERF, V.FN_EXP, V.FN_FAST_EXP, V.FN_SPLINE_EXP)
    rates = {name: renderer.approx_rate(fn) for fn, name in zip(ids, APPROX_FNS)}
    for name, r in rates.items():
        print(f"{name}: {r:.3e} values/s")
        assert 1e11 < r < 4e13, (name, r)
    assert rates["fast_exp"] > rates["spline_exp"]  # two FMA-pipe ops and a conversion against an 18-knot search
    with pytest.raises(V.VrtCudaError):
        renderer.approx_rate(42)
