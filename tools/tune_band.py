#!/usr/bin/env python
"""tools/tune_band.py -- A/B of the banded kernel's tuning knobs on BASELINE configs 5 and 4 (run on the GPU box)."""
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
out = {}
for name, scene in (("config5", pkg.scenes.config5()), ("config4", pkg.scenes.config4())):
    W = 4096
    cam, origin = V.camera_t.app(W, W)
    r.set_gaussians(scene)
    f = r.frame(cam.view_matrix, origin, W, W, (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND, (256, 256), 6.0)
    for ctas in (4, 5):
        r.set_band_tuning(ctas)
        ms = []
        for _ in range(4):
            _, _, st = r.frame_render(f, True, False)
            ms.append((st["ms_tile"], st["ms_render"]))
        out[f"{name}_ctas{ctas}"] = {"ms_tile": float(np.median([m[0] for m in ms[1:]])), "ms_render": float(np.median([m[1] for m in ms[1:]])),
                                    "terms_evaluated": st["terms_executed"]}
    r.set_band_tuning(4)
print(json.dumps(out, indent=1))
r.close()
