"""tools/stress_4m.py -- 4M synthetic Gaussians (BASELINE config 5's sigma range, four times its density) at 4096^2 on one B200:
every cell's list is 250..600 entries, so the whole frame runs on k2_band_long.  Prints the default frame, the round-1 route of
such lists (VRT_CUDA_LONG_BAND=0) and the evaluation of every term, and compares the three pictures."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
V = pkg.vrt
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
W = 4096
scene = pkg.scenes.synthetic(n, 43, -2.6, -2.0)
cam, origin = V.camera_t.app(W, W)
FLAGS = (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND
out, imgs = {}, {}
for key, env, fl in (("k2_band_long", "1", 0), ("k2_render<WIN>", "0", 0), ("every term", "1", V.EVAL_ALL)):
    os.environ["VRT_CUDA_LONG_BAND"] = env
    r = V.Renderer(0)
    r.set_gaussians(scene)
    f = r.frame(cam.view_matrix, origin, W, W, FLAGS | fl, (256, 256), 6.0)
    ms = []
    for k in range(3):
        img, _, st = r.frame_render(f, True, False)
        if k:
            ms.append((st["ms_render"], st["ms_tile"]))
    imgs[key] = img
    out[key] = {"ms_render": round(min(m[0] for m in ms), 2), "ms_tile": round(min(m[1] for m in ms), 2), "listed": st["terms_listed"], "evaluated": st["terms_executed"],
                "saturated": st["terms_saturated"], "terminated": st["terms_terminated"], "max_list": st["max_list"], "list_entries": st["list_entries"], "slice": st["slice"]}
    r.close()
ref = imgs["every term"].view(np.uint8).astype(np.int16)
for key in ("k2_band_long", "k2_render<WIN>"):
    out[key]["max_lsb_vs_every_term"] = int(np.abs(imgs[key].view(np.uint8).astype(np.int16) - ref).max())
print(json.dumps({"n": n, "width": W, **out}))
