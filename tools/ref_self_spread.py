#!/usr/bin/env python
"""tools/ref_self_spread.py -- how far the COMPILED reference is from itself on BASELINE configs 4 and 5 (CPU only).

The north star's tolerance (max abs <= 1e-3 against "the reference's scalar path with the same erf variant") presumes
that this path is reproducible to 1e-3.  On the small-sigma scenes of configs 4/5 it is not: rt.h:40-44 forms
d^2 = |oc|^2 - mu_bar^2 in fp32, which cancels ~|oc|^2 / sigma^2 (1e6 .. 4e6 for sigma = 0.0025 .. 0.005 at depth 5) and
assumes |n| = 1 exactly.  This script measures, with the reference's own object code (oracle/_ref, both ISA builds),
on the same 160 random pixels and lists tests/test_gpu_parity.py::test_config{4,5}_full_size use:

    ref_v4 vs ref_v3          the same source, -march=x86-64-v4 vs -v3  (radiance<transmittance<expf, A&S>>, rt.h:146-164)
    ref_v* vs arbiter         against the closed form in double on exactly-unit rays (oracle/vrt_oracle.c)
    ref(dirs) vs ref(dirs')   the same build fed ray directions that differ by one fp32 rounding of the normalisation
                              (rsqrt-based vec4f_t::normalize vs IEEE sqrt + divide)

Output: one JSON object per config on stdout; tests/golden/ref_self_spread.json is this script's committed output and
tests/test_oracle.py re-measures a subset live.  TEST INFRASTRUCTURE (reads oracle/_ref, never the product).
"""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import cpu_lists  # noqa: E402
from oracle_lib import Oracle, _cpu_flags  # noqa: E402

vp, c_u64, c_i = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int


def load_ref(name):
    p = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(p):
        return None
    L = ctypes.CDLL(p)
    L.ref_radiance.argtypes = [vp, c_u64, vp, vp, c_u64, c_i, vp]
    return L


def ref_radiance(L, lst, origin, dirs, variant):
    g, o, d = (np.ascontiguousarray(a, np.float32) for a in (lst, origin, dirs))
    out = np.zeros((len(d), 4), np.float32)
    L.ref_radiance(g.ctypes.data_as(vp), len(g), o.ctypes.data_as(vp), d.ctypes.data_as(vp), len(d), variant, out.ctypes.data_as(vp))
    return out


def sample_lists(scene, view, origin, W, tiles, n_pix, seed, near_sigmas=12.0, centre=None):
    """The pixels and lists of tests/test_gpu_parity.py::_full_size_case: reference tile predicate AND 12 sigma of the ray.
    centre = (lo, hi): sample rows and columns from [lo, hi) only (scenes that cover part of the frame)."""
    rng = np.random.default_rng(seed)
    if centre is None:
        pix = np.unique(rng.integers(0, W * W, n_pix).astype(np.uint64))
    else:
        pix = np.unique((rng.integers(centre[0], centre[1], n_pix) * W + rng.integers(centre[0], centre[1], n_pix)).astype(np.uint64))
    dirs = Oracle.pixel_dirs(view, origin, W, W, pix)
    mx, my, sg, valid = cpu_lists.projected(scene, view)
    cxs, tw = cpu_lists.tile_centres(tiles)
    tile_px = W // tiles
    lists = []
    for k, p in enumerate(pix):
        row, col = int(p) // W, int(p) % W
        near = np.nonzero(cpu_lists.ray_distance_sigmas(scene, origin, dirs[k : k + 1])[:, 0] < near_sigmas)[0]
        member = cpu_lists.reference_member(mx[near], my[near], sg[near], valid[near], cxs[col // tile_px], cxs[row // tile_px], tw, tw)
        lists.append(near[member])
    return pix, dirs, lists


def measure(pkg, name, scene, W, tiles, n_pix, seed, centre=None):
    cam, origin = pkg.vrt.camera_t.app(W, W)
    pix, dirs, lists = sample_lists(scene, cam.view_matrix, origin, W, tiles, n_pix, seed, centre=centre)
    flags = _cpu_flags()
    libs = {"v3": load_ref("libvrt_ref_v3.so")}
    if {"avx512f", "avx512bw", "avx512dq", "avx512vl", "avx512cd"} <= flags:
        libs["v4"] = load_ref("libvrt_ref_v4.so")
    libs = {k: v for k, v in libs.items() if v is not None}
    # the same directions with the normalisation rounded the IEEE way (sqrt + divide in fp32 of the fp64-exact direction)
    d64 = dirs.astype(np.float64)
    d64[:, :3] /= np.linalg.norm(d64[:, :3], axis=1, keepdims=True)
    dirs_alt = d64.astype(np.float32)
    n = len(pix)
    ideal = np.zeros((n, 4))
    port32 = np.zeros((n, 4), np.float32)
    out = {k: np.zeros((n, 4), np.float32) for k in libs}
    out_alt = {k: np.zeros((n, 4), np.float32) for k in libs}
    for k in range(n):
        lst = scene[lists[k]]
        if not len(lst):
            continue
        ideal[k] = Oracle.radiance(lst, origin, dirs[k : k + 1], 1, "unit")[0]
        port32[k] = Oracle.radiance(lst, origin, dirs[k : k + 1], 1)[0]
        for key, L in libs.items():
            out[key][k] = ref_radiance(L, lst, origin, dirs[k : k + 1], 1)[0]
            out_alt[key][k] = ref_radiance(L, lst, origin, dirs_alt[k : k + 1], 1)[0]
    mx = lambda a, b: float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())
    res = {"config": name, "pixels": int(n), "mean_list": float(np.mean([len(l) for l in lists])), "peak_radiance": float(ideal.max()),
           "max_ulp_between_the_two_direction_roundings": float(np.abs(dirs_alt - dirs).max() / np.spacing(np.float32(1.0)))}
    for key in libs:
        res[f"ref_{key}_vs_arbiter"] = mx(out[key], ideal)
        res[f"ref_{key}_dirs_vs_rounded_dirs"] = mx(out[key], out_alt[key])
    if len(libs) == 2:
        res["ref_v4_vs_ref_v3"] = mx(out["v4"], out["v3"])
    res["restatement_fp32_vs_arbiter"] = mx(port32, ideal)
    for key in libs:
        res[f"restatement_fp32_vs_ref_{key}"] = mx(port32, out[key])
    return res


def main():
    import __graft_entry__ as ge

    pkg = ge.load_package()
    which = sys.argv[1:] or ["4", "5"]
    results = []
    if "4" in which:
        results.append(measure(pkg, "config4: 100k Gaussians @4096^2", pkg.scenes.config4(), 4096, 256, 160, 4))
    if "5" in which:
        results.append(measure(pkg, "config5: 1M Gaussians @4096^2", pkg.scenes.config5(), 4096, 256, 160, 5))
    if "3" in which:  # a well-conditioned control: 64x64 grid, sigma 1/128 at depth 5
        results.append(measure(pkg, "control: 64x64 grid @2048^2", pkg.scenes.grid(64), 2048, 16, 40, 3, centre=(830, 1218)))
    for r in results:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
