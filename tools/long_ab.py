"""A/B of the long-list route on one B200: k2_band_long (default, and with other thresholds of K1's wide-band mark: VRT_CUDA_LONG_WIDE) against k2_render<WIN> (VRT_CUDA_LONG_BAND=0, round 1's
route of lists beyond k2_band's cache) on the bundled OBJ scenes and a dense synthetic frame; prints one JSON line per case.
Usage: python tools/long_ab.py [frames [thresholds, comma-separated]]"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
V = pkg.vrt
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 5
THRESHOLDS = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0.3", "0.6", "0.8"]
FLAGS = (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND


def obj(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}_gaussians.npy")).astype(np.float32)
    return g


def dense(n, mag):
    g = pkg.scenes.synthetic(n, 11, -1.9, -1.5)
    g[:, 9] *= mag
    return g


def dense2(n):
    return pkg.scenes.synthetic(n, 12, -1.6, -1.2)


CASES = [("teapot 1024^2", lambda: obj("teapot"), 1024, 16), ("teapot 2048^2", lambda: obj("teapot"), 2048, 16), ("cube 1024^2", lambda: obj("cube"), 1024, 16),
         ("monkey 1024^2", lambda: obj("monkey"), 1024, 16), ("200k synthetic sigma .012-.03 @1024^2", lambda: dense(200_000, 1.0), 1024, 16),
         ("200k synthetic x20 magnitude @1024^2", lambda: dense(200_000, 20.0), 1024, 16),
         ("100k synthetic sigma .025-.06 @1024^2", lambda: dense2(100_000), 1024, 16)]
for name, make, W, tiles in CASES:
    scene = make()
    if name.startswith(("teapot", "cube", "monkey")):
        scene = scene.copy()
        scene[:, 8] = np.where(scene[:, 8] > 0, scene[:, 8], 0.05)  # the app's default sigma for OBJ vertices
    cam, origin = V.camera_t.app(W, W)
    row = {"case": name, "n": int(len(scene))}
    for key, env, frac in [("long", "1", None)] + [(f"long@{t}", "1", t) for t in THRESHOLDS] + [("long@all", "1", "2"), ("win", "0", None), ("all", "1", None)]:
        os.environ["VRT_CUDA_LONG_BAND"] = env
        os.environ.pop("VRT_CUDA_LONG_WIDE", None)
        if frac:
            os.environ["VRT_CUDA_LONG_WIDE"] = frac
        r = V.Renderer(0)
        r.set_gaussians(scene)
        fl = FLAGS | (V.EVAL_ALL if key == "all" else 0)
        f = r.frame(cam.view_matrix, origin, W, W, fl, (tiles, tiles), 6.0)
        ms, img = [], None
        for k in range(frames + 1):
            img, _, st = r.frame_render(f, True, False)
            if k:
                ms.append((st["ms_render"], st["ms_tile"]))
        ms.sort()
        mr, mt = ms[len(ms) // 2]
        row[key] = {"ms_render": round(mr, 3), "ms_tile": round(mt, 3), "exec": st["terms_executed"], "sat": st["terms_saturated"], "term": st["terms_terminated"],
                    "listed": st["terms_listed"], "max_list": st["max_list"], "slice": st["slice"], "launches": st["n_launches"], "crc": zlib.crc32(img.tobytes())}
        if key != "long":
            d = np.abs(img.view(np.uint8).astype(np.int16) - ref_img.view(np.uint8).astype(np.int16))
            row[key]["max_lsb_vs_long"] = int(d.max())
        else:
            ref_img = img
        r.close()
    print(json.dumps(row), flush=True)
