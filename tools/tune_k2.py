"""Times the render kernel's variants (emitter block Q, packed f32x2 on/off, erf variant) on two workloads.
Run on the GPU box:  python tools/tune_k2.py [--quick]   (writes gpurun_out/tune_k2.json)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    pkg = ge.load_package()
    V = pkg.vrt
    r = V.Renderer(0)
    results = []
    # (a) brute force: 64x64 grid, 2048^2, 16 reference tiles, literal reference lists, no skipping, a 32-row band
    # (b) bounded lists: 100k synthetic Gaussians at 2048^2
    work = [
        ("grid64_ref_noskip_band32", pkg.scenes.grid(64), 2048, (V.MODE8 | V.NO_SKIP), (16, 16), (1024, 1056)),
        ("grid64_ref_skip_band256", pkg.scenes.grid(64), 2048, V.MODE8, (16, 16), (896, 1152)),
        ("synth100k_bound_2048", pkg.scenes.config4(), 2048, (V.MODE4 & ~V.LIST_MASK) | V.LIST_BOUND, (1, 1), (0, 0)),
    ]
    for name, scene, W, flags, tiles, rows in work:
        cam, origin = V.camera_t.app(W, W)
        r.set_gaussians(scene)
        for erf in (V.ERF_AS, V.ERF_EXACT):
            fl = (flags & ~1) | erf
            f = r.frame(cam.view_matrix, origin, W, W, fl, tiles, rows=rows)
            t0 = time.time()
            r.tile(f)
            t_tile = time.time() - t0
            for q in (2, 4, 6, 8):
                for pack in (0, 1):
                    if quick and (q, pack) not in ((4, 1), (4, 0), (8, 1)):
                        continue
                    r.set_tuning(q, pack)
                    best = None
                    for rep in range(3):
                        _, _, st = r.render(f, True, False)
                        if best is None or st["ms_render"] < best["ms_render"]:
                            best = st
                    rate = best["terms_executed"] / (best["ms_render"] * 1e-3)
                    rec = dict(work=name, erf=int(erf), q=q, pack=pack, ms_render=best["ms_render"], ms_tile=best["ms_tile"],
                               terms_executed=best["terms_executed"], terms_listed=best["terms_listed"], rate=rate,
                               max_list=best["max_list"], list_entries=best["list_entries"], host_tile_s=t_tile)
                    results.append(rec)
                    print(json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tune_k2.json"), "w") as fjs:
        json.dump(results, fjs, indent=1)


if __name__ == "__main__":
    main()
