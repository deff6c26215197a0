"""tools/band_k1_profile.py -- the tile call (K0 + K1) of ONE row band of BASELINE config 5 on one GPU, as a rank of an 8-GPU run
sees it (rows 1472..2048 of 4096), for an ncu launch list: which kernels of the tile call do not shrink with the band."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
V = pkg.vrt
W = 4096
scene = pkg.scenes.config5()
cam, origin = V.camera_t.app(W, W)
r = V.Renderer(0)
r.set_gaussians(scene)
rows = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1472, 2048)
f = r.frame(cam.view_matrix, origin, W, W, (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND, (256, 256), 6.0, rows=rows)
for _ in range(4):
    r.tile(f)
    _, _, st = r.render(f, True, False)
print(rows, {k: st[k] for k in ("ms_tile", "ms_render", "n_launches")})
r.close()
