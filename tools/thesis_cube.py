"""The reference's own published benchmark (thesis/main.tex:1797, runtimes.sh: cube.obj, 256x256, --tiles 16, mode 8) on the GPU.
Run on the GPU box:  python tools/thesis_cube.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
V = pkg.vrt
r = V.Renderer(0)
out = []
for name in ("cube", "teapot"):
    scene = np.load(os.path.join(ROOT, "tests", "golden", f"{name}_gaussians.npy"))
    cam, origin = V.camera_t.app(256, 256)
    r.set_gaussians(scene)
    for label, flags in (("mode 8, literal reference lists", V.MODE8), ("mode 8, reference AND 6-sigma lists", (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND),
                         ("same + depth window", (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND | V.DEPTH_WINDOW)):
        f = r.frame(cam.view_matrix, origin, 256, 256, flags, (16, 16), 6.0)
        best = None
        for _ in range(4):
            _, _, st = r.frame_render(f, True, False)
            if best is None or st["ms_tile"] + st["ms_render"] < best["ms_tile"] + best["ms_render"]:
                best = st
        rec = dict(scene=name, mode=label, ms_tile=best["ms_tile"], ms_render=best["ms_render"], terms_listed=best["terms_listed"],
                   terms_executed=best["terms_executed"], evals_per_s=best["terms_executed"] / ((best["ms_tile"] + best["ms_render"]) * 1e-3))
        out.append(rec)
        print(json.dumps(rec))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "thesis_cube.json"), "w"), indent=1)
