import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
pkg = ge.load_package(); V = pkg.vrt
scene = pkg.scenes.synthetic(3000, 7, -1.9, -1.3)
cam, origin = V.camera_t.app(256, 256)
r = V.Renderer(0); r.set_gaussians(scene)
out = {}
for name, flags in (("all_as", V.MODE4), ("bound_as", (V.MODE4 & ~V.LIST_MASK) | V.LIST_BOUND), ("all_exact", V.MODE1), ("all_as_noskip", V.MODE4 | V.NO_SKIP)):
    for q, p in ((4, 1), (2, 0)):
        r.set_tuning(q, p)
        f = r.frame(cam.view_matrix, origin, 256, 256, flags)
        _, rad, st = r.frame_render(f, False, True)
        out[f"{name}_q{q}p{p}"] = rad
        print(name, q, p, st["terms_executed"], st["ms_render"])
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "dump_synth3000.npz"), **out)
