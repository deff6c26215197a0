"""Shared helpers of the parity tests: evaluate the oracle (or the compiled reference) on the pixels of a frame."""
import numpy as np
from oracle_lib import Oracle, Ref


def psnr(a, b, peak=1.0):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 200.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)


def reference_lists(scene, view, tiles):
    """Per-tile index lists of vrt::tile_gaussians (restated by the oracle): list of arrays, row-major, y outer."""
    tw = np.float32(2.0) / np.float32(tiles)
    w, h, counts, idx = Oracle.tile_membership(tw, tw, scene, view)
    assert len(counts) == tiles * tiles == w * h, (len(counts), w, h, tiles)
    offs = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)])
    return [idx[offs[t] : offs[t + 1]] for t in range(tiles * tiles)]


def near_rays(scene, origin, dirs, k_sigma):
    """Indices of the Gaussians within k_sigma standard deviations of ANY of the rays `dirs` (n,4) from `origin`.
    Dropping the rest changes nothing representable: their weight is < exp(-k_sigma^2/2)."""
    oc = scene[:, 4:7].astype(np.float64) - np.asarray(origin, np.float64)[:3]
    n = np.asarray(dirs, np.float64)[:, :3]
    n = n / np.linalg.norm(n, axis=1, keepdims=True)
    mu = oc @ n.T  # (N, rays)
    d2 = np.maximum((oc**2).sum(1)[:, None] - mu**2, 0.0)
    keep = (d2 < (k_sigma * scene[:, 8:9].astype(np.float64)) ** 2).any(1)
    return np.nonzero(keep)[0]


def oracle_radiance(scene, view, origin, W, H, pix, variant, tiles=None, lists=None, f64=False, use_ref=False, near_sigmas=None):
    """Radiance (len(pix), 4) of the reference's scalar path radiance<transmittance<expf, ERF>> at pixel ids `pix`.
    tiles=None: every Gaussian for every pixel (untiled modes); else the pixel's reference-tile list (or `lists`).
    use_ref=True evaluates the compiled reference (oracle/_ref) instead of the restatement."""
    pix = np.asarray(pix, np.uint64)
    dirs = Oracle.pixel_dirs(view, origin, W, H, pix)
    out = np.zeros((len(pix), 4), np.float64 if f64 else np.float32)

    def ev(lst, d):
        if use_ref:
            return Ref.radiance(lst, origin, d, variant)
        return Oracle.radiance(lst, origin, d, variant, f64)

    if tiles is None:
        if near_sigmas is None:
            out[:] = ev(scene, dirs)
        else:  # per-ray prefilter (keeps the O(n^2) oracle affordable on big scenes)
            for k in range(len(pix)):
                out[k] = ev(scene[near_rays(scene, origin, dirs[k : k + 1], near_sigmas)], dirs[k : k + 1])[0]
        return out
    if lists is None:
        lists = reference_lists(scene, view, tiles)
    tw_px, th_px = W // tiles, H // tiles
    rows, cols = (pix // W).astype(np.int64), (pix % W).astype(np.int64)
    tids = (rows // th_px) * tiles + cols // tw_px
    for t in np.unique(tids):
        sel = np.nonzero(tids == t)[0]
        ids = np.asarray(lists[t], np.int64)
        if near_sigmas is not None and len(ids):
            ids = ids[np.isin(ids, near_rays(scene, origin, dirs[sel], near_sigmas))]
        if len(ids) == 0:
            continue
        out[sel] = ev(scene[ids], dirs[sel])
    return out


def pack_image(rad, round_nearest, alpha_quirk):
    """Vectorised orc_pack_pixel over an (h, w, 4) radiance array."""
    v = np.minimum(np.asarray(rad, np.float32), np.float32(1.0)) * np.float32(255.0)
    q = np.rint(v).astype(np.int64) if round_nearest else v.astype(np.int64)  # np.rint = nearest even
    q = q.astype(np.uint32)
    a = (q[..., 3] << 24) if alpha_quirk else np.uint32(0xFF000000)
    return (a | (q[..., 0] << 16) | (q[..., 1] << 8) | q[..., 2]).astype(np.uint32)


def channel_diff_lsb(img_a, img_b):
    """max per-channel difference (in 8-bit steps) between two packed images."""
    d = 0
    for sh in (0, 8, 16, 24):
        a = ((img_a >> sh) & 0xFF).astype(np.int64)
        b = ((img_b >> sh) & 0xFF).astype(np.int64)
        d = max(d, int(np.abs(a - b).max()))
    return d


# ---- alternative approximations (src/vrt/approx.h:10-46) -------------------------------------------------------------
# (abs, rel) tolerance of one table value against the reference's scalar function: a -ffast-math host build, the strict-IEEE
# restatement and the device (FMA contraction, MUFU.RCP / MUFU.EX2) agree only to rounding.  fast_exp turns the float
# a x + b into the result's bit pattern, so one rounding of a x (magnitude up to 2^31, ulp 2^7) is 2^7 / 2^23 = 1.5e-5 relative.
APPROX_TOL = {
    "spline_erf": (2e-6, 0.0),
    "spline_erf_mirror": (2e-6, 0.0),
    "taylor_erf": (1e-6, 0.0),
    "abramowitz_stegun_erf": (2e-6, 0.0),
    "erff": (5e-7, 0.0),
    "expf": (1e-37, 4e-6),
    "fast_exp": (1e-37, 3.2e-5),
    "spline_exp": (1e-6, 0.0),
}


def table_mismatch(name, got, want, x):
    """Largest violation of APPROX_TOL[name] (<= 0 means within tolerance).  Inputs within one ulp of a spline knot or of
    Taylor's +-2 cut may legitimately fall on either side of a discontinuity and are judged against both neighbours."""
    a, r = APPROX_TOL[name]
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float((np.abs(got - want) - (a + r * np.abs(want))).max())
