"""Small frames through the C ABI against the oracle: the heavier code paths at sizes that take milliseconds on the GPU and
seconds under the SIMT interpreter of tests/emu (tests/test_emu.py runs this file there too, VRT_EMU=1).

They reach what the quick parity tests do not: the depth-window kernel and its long-list fall-back, cells split into emitter
slices (+ the combine pass) whose bands must compose bit-exactly, lists walked as contiguous record ranges through the
bulk-copy / mbarrier staging with several chunks in flight, the register-block / packing variants, and a randomised sweep
over image sizes, list lengths around the 32-record staging chunks, list modes, flags and row bands (tests/emu/fuzz_frames.py).
"""
import os
import sys

import numpy as np
import pytest
from parity_util import channel_diff_lsb, oracle_radiance, reference_lists
from test_gpu_parity import all_pixels, check, gpu_at

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bound_flags(V, erf=0):
    return (((V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND) & ~1) | erf


@pytest.mark.parametrize("erf", [0, 1])
def test_depth_window_small(pkg, renderer, erf):
    """The banded default (k2_band) reproduces the evaluation of every term (VRT_CUDA_EVAL_ALL) and resolves most terms by
    saturation; every listed term is evaluated, saturated or dropped by the early exit."""
    V = pkg.vrt
    W = 96
    scene = pkg.scenes.synthetic(2500, 21, -1.9, -1.4)
    cam, origin = V.camera_t.app(W, W, rotation=12.0)
    renderer.set_gaussians(scene)
    f0 = renderer.frame(cam.view_matrix, origin, W, W, bound_flags(V, erf) | V.EVAL_ALL, (12, 12), 6.0)
    img0, rad0, st0 = renderer.frame_render(f0, True, True)
    f1 = renderer.frame(cam.view_matrix, origin, W, W, bound_flags(V, erf), (12, 12), 6.0)
    img1, rad1, st1 = renderer.frame_render(f1, True, True)
    assert float(np.abs(rad0 - rad1).max()) <= 2e-5
    assert channel_diff_lsb(img0, img1) <= 1
    assert st1["terms_listed"] == st0["terms_listed"] and st0["terms_saturated"] == 0 and st0["terms_terminated"] == 0
    resolved = st1["terms_executed"] + st1["terms_saturated"] + st1["terms_terminated"]
    assert abs(resolved - st0["terms_executed"]) <= 1e-2 * st0["terms_executed"]
    assert 0 < st1["terms_saturated"] and st1["terms_executed"] < st0["terms_executed"]
    # the explicit round-1 flag is the same path; without the early exit the image may differ by at most its eps per channel
    f2 = renderer.frame(cam.view_matrix, origin, W, W, bound_flags(V, erf) | V.DEPTH_WINDOW | V.NO_TERMINATE, (12, 12), 6.0)
    _, rad2, st2 = renderer.frame_render(f2, False, True)
    assert st2["terms_terminated"] == 0 and float(np.abs(rad2 - rad1).max()) <= 2e-6
    pix = all_pixels(W, W, 53)
    ideal = oracle_radiance(scene, cam.view_matrix, origin, W, W, pix, 1 - erf, f64="unit", near_sigmas=12)
    check(gpu_at(rad1, pix, W), ideal, f"depth-window mode vs arbiter (erf {erf})")


def test_early_exit_on_an_opaque_scene(pkg, renderer):
    """Transmittance early exit (north_star subsystem 2): behind an opaque layer T is below eps for every pixel of
    a cell, the warp breaks out of the emitter loop, and the image changes by less than the stated bound (1e-6 per channel)
    against the same kernel without the exit -- and stays within tolerance of the arbiter.  A negative magnitude makes
    T non-monotone: K0 flags the scene and the exit must stay off."""
    V = pkg.vrt
    W = 64
    scene = pkg.scenes.synthetic(700, 33, -0.8, -0.5)  # sigma ~1-2 pixels at this size: every pixel is covered several times
    scene[:, 9] *= 25.0  # optical depth 5 .. 37 through every Gaussian's centre: the cloud is opaque after a few of them
    cam, origin = V.camera_t.app(W, W)
    renderer.set_gaussians(scene)
    f_exit = renderer.frame(cam.view_matrix, origin, W, W, bound_flags(V), (8, 8), 6.0)
    f_full = renderer.frame(cam.view_matrix, origin, W, W, bound_flags(V) | V.NO_TERMINATE, (8, 8), 6.0)
    f_all = renderer.frame(cam.view_matrix, origin, W, W, bound_flags(V) | V.EVAL_ALL, (8, 8), 6.0)
    _, rad_exit, st_exit = renderer.frame_render(f_exit, False, True)
    _, rad_full, st_full = renderer.frame_render(f_full, False, True)
    _, rad_all, st_all = renderer.frame_render(f_all, False, True)
    print(f"evaluated {st_exit['terms_executed']:.3e} + saturated {st_exit['terms_saturated']:.3e} + terminated {st_exit['terms_terminated']:.3e} "
          f"of {st_exit['terms_listed']:.3e}; without the exit {st_full['terms_executed']:.3e} evaluated")
    assert st_exit["terms_terminated"] > 0.03 * st_exit["terms_listed"] and st_full["terms_terminated"] == 0
    assert st_exit["terms_executed"] < st_full["terms_executed"]
    assert float(np.abs(rad_exit - rad_full).max()) <= 1e-6 + 1e-6 * float(np.abs(rad_full).max())
    # (optical depths of several hundred per pixel: the two evaluation orders differ by sum(A) * 2^-23 in ln T)
    assert float(np.abs(rad_exit - rad_all).max()) <= 1e-4 * max(1.0, float(rad_all.max()))
    pix = all_pixels(W, W, 29)
    ideal = oracle_radiance(scene, cam.view_matrix, origin, W, W, pix, 1, f64="unit", near_sigmas=12)
    check(gpu_at(rad_exit, pix, W), ideal, "early exit vs arbiter")
    # one absorbing-negative Gaussian: no early exit anywhere in the frame
    bad = scene.copy()
    bad[7, 9] = -1.0
    renderer.set_gaussians(bad)
    _, _, st_bad = renderer.frame_render(f_exit, False, True)
    assert st_bad["terms_terminated"] == 0


def test_depth_window_long_lists_take_the_in_loop_test(pkg, renderer):
    """Lists longer than the banded kernel's per-warp cache (152 entries) go to k2_render's in-loop saturation test; a frame
    that mixes both kinds must still be the image of the full evaluation."""
    V = pkg.vrt
    W = 32
    scene = pkg.scenes.synthetic(2000, 9, -1.0, -0.7)  # wide Gaussians: every cell lists hundreds
    cam, origin = V.camera_t.app(W, W)
    renderer.set_gaussians(scene)
    f0 = renderer.frame(cam.view_matrix, origin, W, W, bound_flags(V) | V.EVAL_ALL, (4, 4), 6.0)
    _, rad0, st0 = renderer.frame_render(f0, False, True)
    assert st0["max_list"] > 160
    f1 = renderer.frame(cam.view_matrix, origin, W, W, bound_flags(V), (4, 4), 6.0)
    _, rad1, st1 = renderer.frame_render(f1, False, True)
    assert float(np.abs(rad0 - rad1).max()) <= 1e-4 * max(1.0, float(rad0.max()))
    assert st1["terms_saturated"] > 0


def _renderer_with_env(V, **env):
    """a context created under the given environment (the long-list knobs are read at vrt_cuda_create)"""
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return V.Renderer(0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


@pytest.mark.parametrize("wide", ["2", "0.4"])
def test_long_lists_share_one_cache_per_cta(pkg, wide):
    """Lists beyond k2_band's per-warp cache (152 < n <= 832) take k2_band_long: one CTA per work item, the list cached once in
    dynamic shared memory, pass A and the emitter blocks dealt to its four warps.  wide = 2: every such list stays in that
    kernel; 0.4 (the default): K1 marks the cells whose band is most of the list and queues them for k2_render<WIN>.  Same image as the evaluation of every term;
    same image as k2_render's in-loop test (VRT_CUDA_LONG_BAND=0, the round-1 route of such lists); split cells
    and bands compose bit-exactly; parity against the arbiter."""
    V = pkg.vrt
    renderer = _renderer_with_env(V, VRT_CUDA_LONG_WIDE=wide)
    try:
        _long_list_checks(pkg, renderer)
    finally:
        renderer.close()


def _long_list_checks(pkg, renderer):
    V = pkg.vrt
    W = 32
    scene = pkg.scenes.synthetic(2000, 9, -1.0, -0.7)  # wide Gaussians: every cell lists hundreds
    cam, origin = V.camera_t.app(W, W)
    renderer.set_gaussians(scene)
    flags = bound_flags(V)
    f_all = renderer.frame(cam.view_matrix, origin, W, W, flags | V.EVAL_ALL, (4, 4), 6.0)
    _, rad_all, st_all = renderer.frame_render(f_all, False, True)
    assert 160 < st_all["max_list"] <= 832
    f = renderer.frame(cam.view_matrix, origin, W, W, flags, (4, 4), 6.0)
    img, rad, st = renderer.frame_render(f, True, True)
    assert float(np.abs(rad - rad_all).max()) <= 1e-4 * max(1.0, float(rad_all.max()))
    resolved = st["terms_executed"] + st["terms_saturated"] + st["terms_terminated"]
    assert abs(resolved - st_all["terms_executed"]) <= 1e-2 * st_all["terms_executed"]
    f_nt = renderer.frame(cam.view_matrix, origin, W, W, flags | V.NO_TERMINATE, (4, 4), 6.0)
    _, rad_nt, st_nt = renderer.frame_render(f_nt, False, True)
    assert st_nt["terms_terminated"] == 0 and float(np.abs(rad_nt - rad).max()) <= 2e-6
    # the round-1 route of the same lists evaluates more of them
    r1 = _renderer_with_env(V, VRT_CUDA_LONG_BAND="0")
    try:
        r1.set_gaussians(scene)
        _, rad_r1, st_r1 = r1.frame_render(r1.frame(cam.view_matrix, origin, W, W, flags, (4, 4), 6.0), False, True)
    finally:
        r1.close()
    print(f"listed {st['terms_listed']:.3e}: k2_band_long evaluates {st['terms_executed']:.3e}, k2_render<WIN> {st_r1['terms_executed']:.3e}")
    # (no ordering of the two counts is asserted here: at 32 x 32 pixels a cell spans a large angle and k2_band's warp-uniform
    # corner-ray bounds are loose, while k2_render<WIN> votes on the lanes' own arguments;
    # tests/test_gpu_parity.py::test_dense_long_lists_take_the_cta_shared_cache has the ratio at a real resolution)
    assert float(np.abs(rad_r1 - rad).max()) <= 1e-4 * max(1.0, float(rad_all.max()))
    # split cells (every slice size) and bands of one slice size
    try:
        for sl in (8, 16, 64, 256):  # (256: no cell of this frame is split -- the kernel stores the pixels itself, no combine pass)
            renderer.set_slice(sl)
            img_s, rad_s, st_s = renderer.frame_render(f, True, True)
            assert st_s["slice"] == sl and float(np.abs(rad_s - rad_all).max()) <= 1e-4 * max(1.0, float(rad_all.max()))
            if sl == 16:
                img2, rad2 = np.zeros_like(img_s), np.zeros_like(rad_s)
                for rows in ((0, 8), (8, 20), (20, 32)):
                    fb = renderer.frame(cam.view_matrix, origin, W, W, flags, (4, 4), 6.0, rows=rows)
                    renderer.tile(fb)
                    renderer.render(fb, True, True, image=img2, radiance=rad2)
                assert np.array_equal(img_s, img2) and np.array_equal(rad_s, rad2)
    finally:
        renderer.set_slice(0)
    pix = all_pixels(W, W, 7)
    ideal = oracle_radiance(scene, cam.view_matrix, origin, W, W, pix, 1, f64="unit", near_sigmas=12)
    check(gpu_at(rad, pix, W), ideal, "k2_band_long vs arbiter")
    # an opaque cloud, four times as deep (an emitter's samples reach 4 sigma in front of it: only items that lie further
    # than that behind the opaque layer can be dropped whole)
    dense = pkg.scenes.synthetic(4000, 9, -1.0, -0.7)
    z = dense[:, 6].copy()
    dense[:, 4] *= (4 * z + 4) / (z + 4)
    dense[:, 5] *= (4 * z + 4) / (z + 4)
    dense[:, 6] = 4 * z
    dense[:, 9] *= 40.0
    renderer.set_gaussians(dense)
    _, rad_e, st_e = renderer.frame_render(f, False, True)
    _, rad_f, st_f = renderer.frame_render(f_nt, False, True)
    print(f"opaque: evaluated {st_e['terms_executed']:.3e} + saturated {st_e['terms_saturated']:.3e} + terminated {st_e['terms_terminated']:.3e} "
          f"of {st_e['terms_listed']:.3e}; without the exit {st_f['terms_executed']:.3e} evaluated")
    assert 152 < st_e["max_list"] <= 832
    assert st_e["terms_terminated"] > 0.2 * st_e["terms_listed"] and st_f["terms_terminated"] == 0
    assert st_e["terms_executed"] < 0.8 * st_f["terms_executed"]
    assert float(np.abs(rad_e - rad_f).max()) <= 1e-6 + 1e-6 * float(np.abs(rad_f).max())


def test_pageable_and_page_locked_host_buffers_give_the_same_frame(pkg, renderer):
    """The host-buffer entries (`vrt_cuda_set_gaussians(host)` + `vrt_cuda_render(host image)`, what the reference's main calls
    with its aligned_malloc image, main.cpp:245): plain pageable memory goes through the context's own page-locked staging
    (threaded host copies; K3 writes the mapped staging), page-locked buffers (the opt-in) are read and written in place.
    Same picture both ways, rows outside the band untouched, also when the scene changes between frames."""
    V = pkg.vrt
    W = 512  # 1 MB of pixels, 27 000 Gaussians = 1.08 MB: above the staging threshold (256 KB) and the registration threshold (1 MB)
    flags = bound_flags(V)
    cam, origin = V.camera_t.app(W, W, rotation=7.0)
    scenes = [pkg.scenes.synthetic(27000, seed, -2.1, -1.8) for seed in (3, 4)]
    frames, keep = {}, []
    for mode in ("pageable", "page-locked"):
        renderer.set_host_pinning(mode == "page-locked")
        try:
            for k, scene in enumerate(scenes):
                if mode == "pageable":
                    # a fresh copy per call, overwritten at once: nothing may refer to the caller's array after set_gaussians returns
                    tmp = scene.copy()
                    renderer.set_gaussians(tmp)
                    tmp[:] = 0
                else:
                    renderer.set_gaussians(scene)  # (registered buffers must outlive the registration: `scenes` does)
                img = np.full((W, W), 0xDEADBEEF, np.uint32)
                keep.append(img)
                f = renderer.frame(cam.view_matrix, origin, W, W, flags, (32, 32), 6.0, rows=(100, 180))
                renderer.tile(f)
                _, _, st = renderer.render(f, True, False, image=img)
                assert np.all(img[:100] == 0xDEADBEEF) and np.all(img[180:] == 0xDEADBEEF) and not np.any(img[100:180] == 0xDEADBEEF)
                frames[(mode, k)] = img
        finally:
            renderer.set_host_pinning(False)
    for k in range(2):
        assert np.array_equal(frames[("pageable", k)], frames[("page-locked", k)])
    assert not np.array_equal(frames[("pageable", 0)], frames[("pageable", 1)])
    # and the frame is the one the device-side entries produce
    renderer.set_gaussians(scenes[1])
    whole, _, _ = renderer.frame_render(renderer.frame(cam.view_matrix, origin, W, W, flags, (32, 32), 6.0, rows=(96, 184)), True, False)
    assert channel_diff_lsb(whole[100:180], frames[("pageable", 1)][100:180]) <= 1


def test_split_cells_and_bands_small(pkg, renderer):
    """Heavy cells split into emitter slices: same image for every slice size (up to fp32 regrouping), bands of one slice
    size compose bit-exactly to the whole frame, parity against the oracle."""
    V = pkg.vrt
    W = 64
    scene = pkg.scenes.synthetic(900, 5, -1.3, -1.0)
    cam, origin = V.camera_t.app(W, W)
    renderer.set_gaussians(scene)
    flags = bound_flags(V)
    f = renderer.frame(cam.view_matrix, origin, W, W, flags, (8, 8))
    try:
        renderer.set_slice(8)
        img, rad, st = renderer.frame_render(f, True, True)
        assert st["slice"] == 8 and st["max_list"] > 24  # lists longer than 3 slices: split cells exist
        img2, rad2 = np.zeros_like(img), np.zeros_like(rad)
        for rows in ((0, 24), (24, 40), (40, 64)):
            fb = renderer.frame(cam.view_matrix, origin, W, W, flags, (8, 8), rows=rows)
            renderer.tile(fb)
            renderer.render(fb, True, True, image=img2, radiance=rad2)
        assert np.array_equal(img, img2) and np.array_equal(rad, rad2)
        for s in (16, 64):
            renderer.set_slice(s)
            _, rad3, st3 = renderer.frame_render(f, False, True)
            assert st3["slice"] == s and float(np.abs(rad3 - rad).max()) <= 2e-5 * max(1.0, float(rad.max()))
    finally:
        renderer.set_slice(0)
    pix = all_pixels(W, W, 29)
    ideal = oracle_radiance(scene, cam.view_matrix, origin, W, W, pix, 1, f64="unit", near_sigmas=12)
    check(gpu_at(rad, pix, W), ideal, "split cells vs arbiter")


def test_contiguous_lists_through_the_bulk_copy_staging(pkg, renderer):
    """NO_SKIP walks the ALL list and caller-supplied tiles_t lists as contiguous record ranges: several 32-record chunks per
    list, double-buffered bulk copies completing on per-warp mbarriers.  Must equal the culled evaluation and the oracle."""
    V = pkg.vrt
    W = 32
    scene = pkg.scenes.synthetic(150, 3, -1.2, -0.9)
    cam, origin = V.camera_t.app(W, W)
    renderer.set_gaussians(scene)
    # untiled: one list of 150 records = 5 chunks
    f_all = renderer.frame(cam.view_matrix, origin, W, W, V.MODE4 | V.NO_SKIP)
    _, rad_all, st_all = renderer.frame_render(f_all, False, True)
    assert st_all["terms_executed"] == st_all["terms_listed"] == 5.0 * 150 * 150 * W * W
    f_cull = renderer.frame(cam.view_matrix, origin, W, W, V.MODE4)
    _, rad_cull, st_cull = renderer.frame_render(f_cull, False, True)
    assert st_cull["terms_executed"] <= st_all["terms_executed"]
    assert float(np.abs(rad_all - rad_cull).max()) <= 2e-5 * max(1.0, float(rad_all.max()))
    pix = all_pixels(W, W, 7)
    ref = oracle_radiance(scene, cam.view_matrix, origin, W, W, pix, 1)
    check(gpu_at(rad_all, pix, W), ref, "ALL list, literal walk (bulk-copy staging)")
    # tiled: the reference's own lists handed over as a tiles_t, walked literally
    lists = reference_lists(scene, cam.view_matrix, 4)
    assert max(len(l) for l in lists) > 32
    f_t = renderer.frame(cam.view_matrix, origin, W, W, V.MODE8 | V.NO_SKIP, (4, 4))
    renderer.set_tile_lists(f_t, [scene[l] for l in lists])
    _, rad_t, st_t = renderer.render(f_t, False, True)
    assert st_t["terms_executed"] == st_t["terms_listed"]
    ref_t = oracle_radiance(scene, cam.view_matrix, origin, W, W, pix, 1, tiles=4, lists=lists)
    check(gpu_at(rad_t, pix, W), ref_t, "tiles_t lists, literal walk (bulk-copy staging)")


def test_register_block_and_packing_variants_small(pkg, renderer):
    """Q = 4 / 8, packed / scalar arithmetic and the exact-erf kernels all evaluate the same sums."""
    V = pkg.vrt
    W = 48
    scene = pkg.scenes.synthetic(400, 11, -1.4, -1.0)
    cam, origin = V.camera_t.app(W, W)
    renderer.set_gaussians(scene)
    for erf in (0, 1):
        f = renderer.frame(cam.view_matrix, origin, W, W, bound_flags(V, erf), (6, 6))
        _, base, _ = renderer.frame_render(f, False, True)
        try:
            for q, p in ((4, 0), (4, 1), (8, 1), (8, 0), (0, 2)):
                renderer.set_tuning(q, p)
                _, r, _ = renderer.frame_render(f, False, True)
                assert float(np.abs(r - base).max()) <= 2e-5 * max(1.0, float(base.max())), (erf, q, p)
        finally:
            renderer.set_tuning(0, 1)
        pix = all_pixels(W, W, 31)
        ideal = oracle_radiance(scene, cam.view_matrix, origin, W, W, pix, 1 - erf, f64="unit", near_sigmas=12)
        check(gpu_at(base, pix, W), ideal, f"bounded lists vs arbiter (erf {erf})")


def test_randomised_small_frames(pkg, renderer):
    """120 random frames (odd image sizes, 0..130 Gaussians, every list mode, NO_SKIP / depth window / pinned slice /
    emitter block, arbitrary row bands): radiance of every band pixel within 1e-3 of the oracle, rows outside the band
    untouched, packed pixels consistent with the radiance (tests/emu/fuzz_frames.py, seed 5)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import fuzz_frames

    rng = np.random.default_rng(5)
    worst = 0.0
    for _ in range(120):
        _, err = fuzz_frames.run_case(pkg, renderer, rng)
        worst = max(worst, err)
    print(f"worst max-abs radiance error over 120 random frames: {worst:.3e}")
    # 40 more with 153..260 Gaussians (lists beyond k2_band's per-warp cache: k2_band_long / k2_render<WIN>), thin bands
    worst = 0.0
    for _ in range(40):
        _, err = fuzz_frames.run_case(pkg, renderer, rng, sizes=fuzz_frames.SIZES_LONG, max_band_rows=5)
        worst = max(worst, err)
    print(f"worst max-abs radiance error over 40 random long-list frames: {worst:.3e}")


def test_k1_bins_cells_and_bands(pkg, renderer):
    """Tile-only: 1024 x 768 (32 x 24 bins of 32 x 32 pixels, 128 x 192 cells), reference tiles that are not square, a row band
    that cuts through bins.  The band's lists must be exactly the full frame's lists for
    the cells it renders and empty elsewhere; sampled cells are conservative against the exact per-ray criterion and no
    longer than the four-plane box (as tests/test_gpu_parity.py checks at 256^2 with a single level pair)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import cpu_lists
    from oracle_lib import Oracle

    V = pkg.vrt
    W, H, k = 1024, 768, 6.0
    scene = pkg.scenes.synthetic(20000, 17, -2.2, -1.5)
    cam, origin = V.camera_t.app(W, H, rotation=-19.0)
    renderer.set_gaussians(scene)
    ncx, ncy = W // 8, H // 4
    for flags, tiles in (((V.MODE4 & ~V.LIST_MASK) | V.LIST_BOUND, (1, 1)), (bound_flags(V), (32, 24))):
        f = renderer.frame(cam.view_matrix, origin, W, H, flags, tiles, k)
        renderer.tile(f)
        counts, idx = renderer.get_lists()
        assert len(counts) == ncx * ncy
        offs = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)])
        rows = (200, 520) if tiles == (1, 1) else (224, 544)  # 544 = 17 tile rows of 32 px
        fb = renderer.frame(cam.view_matrix, origin, W, H, flags, tiles, k, rows=rows)
        renderer.tile(fb)
        bcounts, bidx = renderer.get_lists()
        boffs = np.concatenate([[0], np.cumsum(bcounts, dtype=np.int64)])
        cy = np.arange(ncx * ncy) // ncx
        inside = (cy * 4 + 4 > rows[0]) & (cy * 4 < rows[1])
        assert np.array_equal(bcounts[inside], counts[inside]) and not bcounts[~inside].any()
        first, last = int(np.nonzero(inside)[0][0]), int(np.nonzero(inside)[0][-1])
        assert np.array_equal(bidx, idx[offs[first] : offs[last + 1]]), "a band's lists differ from the full frame's"
        assert boffs[-1] == offs[last + 1] - offs[first]
        if tiles != (1, 1):
            continue  # the per-ray criterion below is the untiled one
        rng = np.random.default_rng(8)
        n_got = n_box = 0
        for cell in rng.integers(0, ncx * ncy, 40):
            cx, cyy = int(cell) % ncx, int(cell) // ncx
            got = set(idx[offs[cell] : offs[cell + 1]].tolist())
            pix = np.array([(cyy * 4 + r) * W + cx * 8 + c for r in range(4) for c in range(8)], np.uint64)
            dirs = Oracle.pixel_dirs(cam.view_matrix, origin, W, H, pix)
            need = set(np.nonzero((cpu_lists.ray_distance_sigmas(scene, origin, dirs) <= k * 0.999).any(1))[0].tolist())
            planes = cpu_lists.rect_planes(cam.view_matrix, origin, W, H, cx * 8, cx * 8 + 8, cyy * 4, cyy * 4 + 4)
            box = set(np.nonzero(cpu_lists.rect_bound_member(scene, origin, planes, k * 1.001))[0].tolist())
            assert need <= got <= box, (cell, sorted(need - got), sorted(got - box))
            n_got, n_box = n_got + len(got), n_box + len(box)
            # the list is in depth order along the cell's centre ray (what the banded kernel walks), ties by Gaussian index
            order = idx[offs[cell] : offs[cell + 1]].astype(np.int64)
            centre = np.array([[cx * 8 + 3.5, cyy * 4 + 1.5]])
            inv = np.linalg.inv(np.asarray(cam.view_matrix, np.float64).reshape(4, 4).T)
            u, v = -1.0 + centre[0, 0] / (W / 2.0), -1.0 + centre[0, 1] / (H / 2.0)
            d = inv[:3, 0] * u + inv[:3, 1] * v + inv[:3, 3] - np.asarray(origin, np.float64)[:3]
            depth = (scene[order, 4:7].astype(np.float64) - np.asarray(origin, np.float64)[:3]) @ (d / np.linalg.norm(d))
            assert np.all(np.diff(depth) >= -1e-5), (cell, depth)
        assert n_got <= n_box


def test_long_lists_are_sorted_in_global_memory(pkg, renderer):
    """Cells whose list exceeds the leaf warp's shared-memory buffer (512 entries) are written unsorted and sorted in place in
    the index array (the same all-ascending bitonic network, keys refetched): the order must still be depth along the cell's
    centre ray with ties by index -- a pure function of the frame -- and two builds must agree entry for entry."""
    V = pkg.vrt
    W, H, k = 32, 24, 6.0
    scene = pkg.scenes.synthetic(2600, 41, -0.7, -0.4)  # wide Gaussians: every cell lists most of the scene
    cam, origin = V.camera_t.app(W, H, rotation=9.0)
    renderer.set_gaussians(scene)
    f = renderer.frame(cam.view_matrix, origin, W, H, (V.MODE4 & ~V.LIST_MASK) | V.LIST_BOUND, (1, 1), k)
    renderer.tile(f)
    counts, idx = renderer.get_lists()
    assert counts.max() > 512 and counts.min() > 128
    renderer.tile(f)
    counts2, idx2 = renderer.get_lists()
    assert np.array_equal(counts, counts2) and np.array_equal(idx, idx2)
    offs = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)])
    inv = np.linalg.inv(np.asarray(cam.view_matrix, np.float64).reshape(4, 4).T)
    o = np.asarray(origin, np.float64)[:3]
    ncx = W // 8
    for cell in (0, 5, len(counts) - 1):
        cx, cy = cell % ncx, cell // ncx
        u, v = -1.0 + (cx * 8 + 3.5) / (W / 2.0), -1.0 + (cy * 4 + 1.5) / (H / 2.0)
        d = inv[:3, 0] * u + inv[:3, 1] * v + inv[:3, 3] - o
        order = idx[offs[cell] : offs[cell + 1]].astype(np.int64)
        assert len(set(order.tolist())) == len(order)
        depth = (scene[order, 4:7].astype(np.float64) - o) @ (d / np.linalg.norm(d))
        assert np.all(np.diff(depth) >= -1e-5), cell
    # and the frame renders (the lists beyond the banded kernel's cache take k2_render's in-loop saturation test)
    _, rad, st = renderer.render(f, False, True)
    pix = all_pixels(W, H, 37)
    ideal = oracle_radiance(scene, cam.view_matrix, origin, W, H, pix, 1, f64="unit", near_sigmas=12)
    check(gpu_at(rad, pix, W), ideal, "long lists vs arbiter")


def test_entry_totals_are_64_bit_and_refused_not_wrapped(pkg):
    """The sum of the list lengths is kept in 64 bits (the leaf cursor, k1_total64 beside the 32-bit scan of the literal lists);
    past the 32-bit offset range the tile call returns VRT_CUDA_E_NOMEM instead of writing through wrapped offsets.
    VRT_CUDA_MAX_LIST_ENTRIES lowers that limit for this test (read once per process: a child process)."""
    import subprocess

    code = (
        "import sys, os, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import __graft_entry__ as ge\n"
        "pkg = ge.load_package()\n"
        "if os.environ.get('VRT_EMU') == '1':\n"
        "    import ctypes; sys.path.insert(0, os.path.join(%r, 'tests', 'emu')); import build_emu\n"
        "    lib = ctypes.CDLL(build_emu.build())\n"
        "    for sym, (res, args) in pkg._ffi.CUDA_SYMBOLS.items():\n"
        "        fn = getattr(lib, sym); fn.restype, fn.argtypes = res, args\n"
        "    pkg._ffi._cuda = lib\n"
        "V = pkg.vrt; r = V.Renderer(0)\n"
        "scene = pkg.scenes.synthetic(600, 3, -1.0, -0.7); cam, origin = V.camera_t.app(64, 64); r.set_gaussians(scene)\n"
        "out = []\n"
        "for flags, tiles in (((V.MODE4 & ~V.LIST_MASK) | V.LIST_BOUND, (1, 1)), (V.MODE8 | V.NO_SKIP, (8, 8))):\n"
        "    try:\n"
        "        r.tile(r.frame(cam.view_matrix, origin, 64, 64, flags, tiles)); out.append('built')\n"
        "    except V.VrtCudaError as e:\n"
        "        out.append('refused' if '= -4' in str(e) and '32-bit' in str(e) else 'other: ' + str(e))\n"
        "print(' | '.join(out))\n"
    ) % (ROOT, ROOT)
    env = dict(os.environ, VRT_CUDA_MAX_LIST_ENTRIES="2000")
    env.pop("LD_PRELOAD", None)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().splitlines()[-1] == "refused | refused", r.stdout
    env.pop("VRT_CUDA_MAX_LIST_ENTRIES")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1] == "built | built", r.stdout + r.stderr[-2000:]
