"""Times the render kernel's variants (emitter block Q, packed f32x2 / occupancy variant, erf variant).
Run on the GPU box:  python tools/tune_k2.py [--quick] [--work NAME]   (writes gpurun_out/tune_k2.json)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    only = sys.argv[sys.argv.index("--work") + 1] if "--work" in sys.argv else None
    pkg = ge.load_package()
    V = pkg.vrt
    r = V.Renderer(0)
    results = []
    bound = (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND
    work = [
        # brute force: 64x64 grid, 2048^2, 16 reference tiles, literal reference lists, no skipping, a 32-row band
        ("grid64_ref_noskip_band32", lambda: pkg.scenes.grid(64), 2048, (V.MODE8 | V.NO_SKIP), (16, 16), (1024, 1056)),
        ("config4_bound", pkg.scenes.config4, 4096, bound, (256, 256), (0, 0)),
        ("config5_bound", pkg.scenes.config5, 4096, bound, (256, 256), (0, 0)),
        ("config5_window", pkg.scenes.config5, 4096, bound | V.DEPTH_WINDOW, (256, 256), (0, 0)),
        ("config4_window", pkg.scenes.config4, 4096, bound | V.DEPTH_WINDOW, (256, 256), (0, 0)),
    ]
    variants = [(8, 1), (4, 1), (8, 0), (4, 0)]
    if quick:
        variants = [(8, 1), (4, 1), (8, 2)]
    for name, scene_fn, W, flags, tiles, rows in work:
        if only and name != only:
            continue
        cam, origin = V.camera_t.app(W, W)
        r.set_gaussians(scene_fn())
        for erf in (V.ERF_AS, V.ERF_EXACT):
            if quick and erf == V.ERF_EXACT:
                continue
            fl = (flags & ~1) | erf
            f = r.frame(cam.view_matrix, origin, W, W, fl, tiles, 6.0, rows=rows)
            t0 = time.time()
            r.tile(f)
            t_tile = time.time() - t0
            for q, pack in variants:
                r.set_tuning(q, pack)
                best = None
                for rep in range(2):
                    st = r.render_device(f, 0, 0, want_stats=True) if False else r.render(f, True, False)[2]
                    if best is None or st["ms_render"] < best["ms_render"]:
                        best = st
                rate = best["terms_executed"] / (best["ms_render"] * 1e-3)
                rec = dict(work=name, erf=int(erf), q=q, pack=pack, ms_render=best["ms_render"], ms_tile=best["ms_tile"],
                           terms_executed=best["terms_executed"], terms_listed=best["terms_listed"], rate=rate,
                           frac=rate * 15 / 74.45e12, terms_saturated=best["terms_saturated"], max_list=best["max_list"], list_entries=best["list_entries"], host_tile_s=t_tile)
                results.append(rec)
                print(json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tune_k2.json"), "w") as fjs:
        json.dump(results, fjs, indent=1)


if __name__ == "__main__":
    main()
