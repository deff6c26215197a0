"""Randomised small frames through the C ABI against the oracle (run by tests/test_gpu_small_frames.py on the GPU and by tests/test_emu.py under the
SIMT interpreter; by hand: `python tests/emu/fuzz_frames.py --cases 200 [--emu] [--seed 1]`).

Every case draws an image size (any width / height, not multiples of the 8x4 cell), a tile count that divides it, a scene
size around the staging boundaries (0, 1, 31, 32, 33, 64, 65, ...), a list mode, an erf variant, optional NO_SKIP /
depth window / pinned slice / emitter block, and a row band (any rows, also not cell-aligned), renders it and compares the
float radiance of every pixel of the band with the oracle's scalar path on the lists the mode defines (tolerance of the
north star: 1e-3 absolute); rows outside the band must stay untouched and the packed image must follow from the radiance.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]

SIZES = [0, 1, 2, 7, 31, 32, 33, 63, 64, 65, 97, 130]
# lists beyond k2_band's per-warp cache (152 entries): k2_band_long, or k2_render<WIN> for the cells K1 marks as wide
SIZES_LONG = [153, 170, 200, 260]


def tile_lists(scene, view, tx, ty):
    """vrt::tile_gaussians(2/tx, 2/ty, scene, view) restated by the oracle: one index array per tile, row-major, y outer"""
    from oracle_lib import Oracle

    w, h, counts, idx = Oracle.tile_membership(np.float32(2.0) / np.float32(tx), np.float32(2.0) / np.float32(ty), scene, view)
    assert (w, h) == (tx, ty) and len(counts) == tx * ty, (w, h, tx, ty)
    offs = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)])
    return [idx[offs[t] : offs[t + 1]] for t in range(tx * ty)]


def run_case(pkg, renderer, rng, verbose=False, sizes=SIZES, max_band_rows=None):
    from parity_util import oracle_radiance, pack_image, channel_diff_lsb

    V = pkg.vrt
    tx, ty = int(rng.choice([1, 1, 2, 3, 4, 5, 8])), int(rng.choice([1, 1, 2, 3, 4, 5, 8]))  # reference tiles per axis, not square
    W, H = tx * int(rng.integers(1, max(2, 72 // tx))), ty * int(rng.integers(1, max(2, 56 // ty)))
    n = int(rng.choice(sizes))
    scene = pkg.scenes.synthetic(n, int(rng.integers(1, 1 << 30)), -1.3, -0.8) if n else np.zeros((0, 10), np.float32)
    cam, origin = V.camera_t.app(W, H, rotation=float(rng.uniform(-40, 40)))
    erf = int(rng.integers(0, 2))
    kind = str(rng.choice(["all", "bound", "reference", "reference_bound", "host_lists"]))
    tiled = kind in ("reference", "reference_bound", "host_lists")
    if not tiled:
        tx = ty = 1
    lm = {"all": V.LIST_ALL, "bound": V.LIST_BOUND, "reference": V.LIST_REFERENCE, "reference_bound": V.LIST_REFERENCE_BOUND, "host_lists": V.LIST_REFERENCE}[kind]
    nearest, alpha_w = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    flags = erf | lm | (V.QUANT_NEAREST if nearest else 0) | (V.ALPHA_FROM_W if alpha_w else 0)
    bounded = kind in ("bound", "reference_bound")
    if not bounded and rng.integers(0, 3) == 0:
        flags |= V.NO_SKIP
    if not (flags & V.NO_SKIP):  # the banded kernel is the default: half of the cases ask for the full evaluation instead
        pick = int(rng.integers(0, 4))
        flags |= (V.EVAL_ALL, V.EVAL_ALL, 0, V.NO_TERMINATE)[pick]
    rows = (0, 0)
    if rng.integers(0, 2) == 0 and H > 1:
        a = int(rng.integers(0, H - 1))
        rows = (a, int(rng.integers(a + 1, H + 1)))
    if max_band_rows is not None and H > max_band_rows:  # (long lists: the oracle's work per pixel grows with n^2)
        a = rows[0] if rows != (0, 0) else int(rng.integers(0, H - max_band_rows + 1))
        rows = (a, min(H, a + max_band_rows, rows[1] if rows != (0, 0) else H))
    slice_ = int(rng.choice([0, 0, 8, 16, 64]))  # (128 and 256 exist too; kept out of this draw so that the seeded sequences stay what they were)
    bound_k = float(rng.choice([0.0, 6.0, 8.0]))  # 0 selects the default (6)
    q = int(rng.choice([0, 0, 4, 8]))
    desc = f"{W}x{H} n={n} {kind} tiles={tx}x{ty} erf={erf} flags={flags:#x} rows={rows} slice={slice_} q={q}"
    if verbose:
        print(desc, flush=True)
    renderer.set_slice(slice_)
    renderer.set_tuning(q, 1)
    try:
        renderer.set_gaussians(scene)
        f = renderer.frame(cam.view_matrix, origin, W, H, flags, (tx, ty), bound_k, rows=rows)
        lists = tile_lists(scene, cam.view_matrix, tx, ty) if tiled else None
        if kind == "host_lists":
            renderer.set_tile_lists(f, [scene[l] for l in lists])
        else:
            renderer.tile(f)
        sentinel = np.float32(-7.0)
        img = np.full((H, W), 0xDEADBEEF, np.uint32)
        rad = np.full((H, W, 4), sentinel, np.float32)
        _, _, st = renderer.render(f, True, True, image=img, radiance=rad)
    finally:
        renderer.set_slice(0)
        renderer.set_tuning(0, 1)
    r0, r1 = rows if rows != (0, 0) else (0, H)
    assert np.all(rad[:r0] == sentinel) and np.all(rad[r1:] == sentinel), desc + ": rows outside the band were written"
    assert np.all(img[:r0] == 0xDEADBEEF) and np.all(img[r1:] == 0xDEADBEEF), desc + ": rows outside the band were written"
    band = rad[r0:r1]
    assert not np.any(band == sentinel), desc + ": a pixel of the band was not written"
    pix = (np.arange(r0, r1, dtype=np.uint64)[:, None] * np.uint64(W) + np.arange(W, dtype=np.uint64)[None, :]).reshape(-1)
    if n == 0:
        ref = np.zeros((len(pix), 4), np.float32)
    elif tiled:
        # per reference tile (tx x ty of them, W/tx x H/ty pixels each): the scalar path on that tile's list
        ref = np.zeros((len(pix), 4), np.float64 if bounded else np.float32)
        tid = ((pix // np.uint64(W)).astype(np.int64) // (H // ty)) * tx + (pix % np.uint64(W)).astype(np.int64) // (W // tx)
        for t in np.unique(tid):
            sel = np.nonzero(tid == t)[0]
            one = lists[int(t)]
            if len(one):
                ref[sel] = oracle_radiance(scene[one], cam.view_matrix, origin, W, H, pix[sel], 1 - erf, f64="unit" if bounded else False,
                                           near_sigmas=12 if bounded else None)
    else:
        ref = oracle_radiance(scene, cam.view_matrix, origin, W, H, pix, 1 - erf, f64="unit" if bounded else False, near_sigmas=12 if bounded else None)
    got = band.reshape(-1, 4)
    finite = np.isfinite(ref).all(1)  # a ray through the camera position is 0/0 in the reference too
    err = float(np.abs(got[finite].astype(np.float64) - ref[finite].astype(np.float64)).max()) if finite.any() else 0.0
    assert err <= 1e-3, f"{desc}: max abs radiance error {err:.3e}"
    want = pack_image(band, nearest, alpha_w)
    ok = np.isfinite(band).all(-1)
    assert channel_diff_lsb(np.where(ok, img[r0:r1], 0), np.where(ok, want, 0)) == 0, desc + ": packed pixels do not follow from the radiance"
    if flags & V.NO_SKIP:
        assert st["terms_executed"] == st["terms_listed"], desc
    return desc, err


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--emu", action="store_true", help="run against the interpreter build (VRT_EMU_LIB or tests/emu/_build)")
    ap.add_argument("--long-cases", type=int, default=0, help="further cases with 153..260 Gaussians and thin bands (lists beyond k2_band's cache)")
    ap.add_argument("-v", action="store_true")
    a = ap.parse_args()
    import __graft_entry__ as ge

    pkg = ge.load_package()
    if a.emu:
        import ctypes

        import build_emu

        lib = ctypes.CDLL(os.environ.get("VRT_EMU_LIB") or build_emu.build())
        for sym, (res, args) in pkg._ffi.CUDA_SYMBOLS.items():
            fn = getattr(lib, sym)
            fn.restype, fn.argtypes = res, args
        pkg._ffi._cuda = lib
    renderer = pkg.vrt.Renderer(0)
    rng = np.random.default_rng(a.seed)
    worst = 0.0
    for k in range(a.cases):
        desc, err = run_case(pkg, renderer, rng, a.v)
        worst = max(worst, err)
    for k in range(a.long_cases):
        desc, err = run_case(pkg, renderer, rng, a.v, sizes=SIZES_LONG, max_band_rows=5)
        worst = max(worst, err)
    print(f"fuzz ok: {a.cases + a.long_cases} cases, worst max-abs radiance error {worst:.3e}")


if __name__ == "__main__":
    main()
