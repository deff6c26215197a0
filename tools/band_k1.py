#!/usr/bin/env python3
"""One rank's share of a multi-GPU frame on a single GPU: tile + render one row band of BASELINE config 5 (for profiling the
per-rank fixed cost of K0/K1; `ncu --metrics gpu__time_duration.sum` of this script lists the band's launches)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
V = pkg.vrt
rows = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1440, 2064)
W, tiles = 4096, 256
scene = pkg.scenes.synthetic(1_000_000, 43, -2.6, -2.0)
cam, origin = V.camera_t.app(W, W)
r = V.Renderer(0)
r.set_gaussians(scene)
flags = (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND
f = r.frame(cam.view_matrix, origin, W, W, flags, (tiles, tiles), 6.0, rows=rows)
for it in range(4):
    t0 = time.time()
    r.tile(f)
    t1 = time.time()
    _, _, st = r.render(f, False, False)
    print(f"band {rows}: ms_tile {st['ms_tile']:.3f} (host wall {1e3 * (t1 - t0):.3f}), ms_render {st['ms_render']:.3f}, launches {st['n_launches']}")
r.close()
