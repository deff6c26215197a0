"""tools/long_profile.py -- a few frames of a dense narrow-band scene whose lists exceed k2_band's per-warp cache (200k synthetic
Gaussians, sigma 0.012-0.03, 1024^2; optionally x20 magnitudes = opaque), for `ncu -k regex:k2_band_long`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
V = pkg.vrt
mag = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
scene = pkg.scenes.synthetic(200_000, 11, -1.9, -1.5)
scene[:, 9] *= mag
W = 1024
cam, origin = V.camera_t.app(W, W)
r = V.Renderer(0)
r.set_gaussians(scene)
f = r.frame(cam.view_matrix, origin, W, W, (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND, (16, 16), 6.0)
for _ in range(3):
    _, _, st = r.frame_render(f, True, False)
print({k: st[k] for k in ("ms_render", "ms_tile", "terms_listed", "terms_executed", "terms_saturated", "terms_terminated", "max_list", "slice", "n_launches")})
r.close()
