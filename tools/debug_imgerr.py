import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
from parity_util import oracle_radiance, reference_lists
pkg = ge.load_package(); V = pkg.vrt
scene = pkg.scenes.img_error_grid(); view = np.eye(4, dtype=np.float32).reshape(16); origin = np.zeros(4, np.float32)
r = V.Renderer(0); r.set_gaussians(scene)
pix = np.arange(0, 65536, 5, dtype=np.uint64)
lists = reference_lists(scene, view, 16)
for mode, variant in (("MODE5", 0), ("MODE8", 1)):
    ref64 = oracle_radiance(scene, view, origin, 256, 256, pix, variant, tiles=16, lists=lists, f64=True)
    ref32 = oracle_radiance(scene, view, origin, 256, 256, pix, variant, tiles=16, lists=lists)
    print(mode, "oracle32 vs 64:", np.nanmax(np.abs(ref32 - ref64)))
    for q, p, extra in ((4, 1, 0), (2, 0, 0), (4, 1, V.NO_SKIP)):
        r.set_tuning(q, p)
        f = r.frame(view, origin, 256, 256, getattr(V, mode) | extra, (16, 16))
        _, rad, st = r.frame_render(f, False, True)
        g = rad.reshape(-1, 4)[pix.astype(np.int64)]
        d = np.abs(g - ref64)
        w = np.unravel_index(np.nanargmax(d), d.shape)
        print(mode, q, p, extra, "max err", np.nanmax(d), "at pix", int(pix[w[0]]), divmod(int(pix[w[0]]), 256), "ch", w[1], "gpu", g[w[0]], "ref", ref64[w[0]], "nan:", np.isnan(g).sum(), np.isnan(ref64).sum())
        big = np.nonzero(d.max(1) > 2e-4)[0]
        print("   n bad", len(big), "rows", sorted(set((pix[big] // 256).tolist()))[:20], "cols", sorted(set((pix[big] % 256).tolist()))[:20])
