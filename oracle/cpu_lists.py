"""oracle/cpu_lists.py -- numpy restatement of the list semantics, TEST INFRASTRUCTURE ONLY.

Used by tests/ (to check K1's bounded lists) and by bench.py's cpu_baseline / --impl reference legs (to hand the
reference's own render entry the lists of a sample of tiles without touching the CUDA path).  Never imported by the
product.

    projected(scene, view)                 the tiling projection of vrt::tile_gaussians (src/vrt/rt.cpp:35-45)
    tile_centres(tiles)                    the float-accumulated tile centres (rt.cpp:47-49)
    reference_member(...)                  the cull predicate of rt.cpp:57-59 for one tile
    rect_bound_member(...)                 "within k sigma of the frustum of a pixel rectangle" (the *_BOUND list modes)
    plane_points(view, W, H, rows, cols)   camera_t::update's projection-plane points (src/vrt/camera.cpp:60-70)
"""
import numpy as np

f32 = np.float32


def inverse_view(view16):
    return np.linalg.inv(np.asarray(view16, np.float64).reshape(4, 4).T)  # column-major in


def projected(scene, view16):
    """-> (mu'x, mu'y, sigma', valid) as float32 arrays: proj = view (mu,1); drop proj.z < 1 and sigma' < 1e-5."""
    v = np.asarray(view16, f32)
    x, y, z = scene[:, 4].astype(f32), scene[:, 5].astype(f32), scene[:, 6].astype(f32)
    px = (v[0] * x + v[4] * y) + (v[8] * z + v[12])
    py = (v[1] * x + v[5] * y) + (v[9] * z + v[13])
    pz = (v[2] * x + v[6] * y) + (v[10] * z + v[14])
    with np.errstate(divide="ignore", invalid="ignore"):
        sg = scene[:, 8].astype(f32) / pz
        mx, my = px / pz, py / pz
    valid = ~(pz < f32(1.0)) & ~(sg < f32(1e-5))
    return mx, my, sg, valid


def tile_centres(tiles):
    t = f32(2.0) / f32(tiles)
    out, x = [], f32(-1.0) + t / f32(2)
    while x < f32(1.0):
        out.append(x)
        x = f32(x + t)
    return np.array(out, f32), t


def reference_member(mx, my, sg, valid, cx, cy, tw, th):
    s33 = f32(3.3) * sg
    return valid & (np.abs(cx - mx) <= np.abs(cx) + tw / f32(2) + s33) & (np.abs(cy - my) <= np.abs(cy) + th / f32(2) + s33)


def rect_planes(view16, origin, W, H, x0, x1, y0, y1):
    """Outward unit normals (4,3) of the side planes of the frustum through pixel samples [x0,x1) x [y0,y1)."""
    inv = inverse_view(view16)
    R, U = inv[:3, 0], inv[:3, 1]
    Wv = inv[:3, 3] - np.asarray(origin, np.float64)[:3]
    u0, u1 = -1.0 + x0 / (W / 2.0), -1.0 + (x1 - 1) / (W / 2.0)
    v0, v1 = -1.0 + y0 / (H / 2.0), -1.0 + (y1 - 1) / (H / 2.0)

    def plane(span_a, span_b, towards, sign):
        n = np.cross(span_a, span_b)
        n /= np.linalg.norm(n)
        return n if np.dot(n, towards) * sign >= 0 else -n

    return np.array([plane(U, u0 * R + Wv, R, -1), plane(U, u1 * R + Wv, R, +1), plane(R, v0 * U + Wv, U, -1), plane(R, v1 * U + Wv, U, +1)])


def rect_bound_member(scene, origin, planes, k_sigma):
    """Gaussians whose centre lies within k sigma of the rectangle's ray frustum or of its mirror image (the closed
    form integrates along the whole line, src/vrt/rt.h:155-160)."""
    oc = scene[:, 4:7].astype(np.float64) - np.asarray(origin, np.float64)[:3]
    d = oc @ planes.T  # (N,4) signed distances, positive = outside
    lim = (k_sigma * scene[:, 8].astype(np.float64))[:, None]
    return (d <= lim).all(1) | (-d <= lim).all(1)


def ray_distance_sigmas(scene, origin, dirs):
    """(N, rays) distance of every centre to every ray's line, in units of its sigma."""
    oc = scene[:, 4:7].astype(np.float64) - np.asarray(origin, np.float64)[:3]
    n = np.asarray(dirs, np.float64)[:, :3]
    n = n / np.linalg.norm(n, axis=1, keepdims=True)
    mu = oc @ n.T
    d2 = np.maximum((oc**2).sum(1)[:, None] - mu**2, 0.0)
    return np.sqrt(d2) / scene[:, 8:9].astype(np.float64)


def plane_points(view16, W, H, rows, cols):
    """Projection-plane points (3, len(rows), len(cols)) float32 as camera_t::update stores them (camera.cpp:60-70)."""
    inv = np.linalg.inv(np.asarray(view16, np.float64).reshape(4, 4).T).astype(f32)
    u = (f32(-1.0) + np.asarray(cols, f32) / f32(W / 2.0)).astype(f32)
    v = (f32(-1.0) + np.asarray(rows, f32) / f32(H / 2.0)).astype(f32)
    out = np.zeros((3, len(v), len(u)), f32)
    for i in range(3):
        out[i] = inv[i, 0] * u[None, :] + inv[i, 1] * v[:, None] + inv[i, 3]
    return out


def sample_tile_lists(scene, view16, origin, W, H, tiles, tile_ids, k_sigma, use_reference=True):
    """Lists (index arrays) of the given reference tiles under "reference predicate AND k-sigma bound of the tile's
    pixel rectangle".  A screen-space prefilter keeps the per-tile tests cheap on million-Gaussian scenes."""
    tile_w, tile_h = W // tiles, H // tiles
    cxs, tw = tile_centres(tiles)
    tids = np.asarray(tile_ids, np.int64)
    txs, tys = tids % tiles, tids // tiles
    mx, my, sg, valid = projected(scene, view16) if use_reference else (None, None, None, None)
    # prefilter: the bound of the bounding rectangle of all requested tiles
    x0, x1 = int(txs.min()) * tile_w, (int(txs.max()) + 1) * tile_w
    y0, y1 = int(tys.min()) * tile_h, (int(tys.max()) + 1) * tile_h
    pre = np.nonzero(rect_bound_member(scene, origin, rect_planes(view16, origin, W, H, x0, x1, y0, y1), k_sigma))[0]
    sub = scene[pre]
    lists = []
    for tx, ty in zip(txs, tys):
        pl = rect_planes(view16, origin, W, H, tx * tile_w, (tx + 1) * tile_w, ty * tile_h, (ty + 1) * tile_h)
        m = rect_bound_member(sub, origin, pl, k_sigma)
        if use_reference:
            m &= reference_member(mx[pre], my[pre], sg[pre], valid[pre], cxs[tx], cxs[ty], tw, tw)
        lists.append(pre[m])
    return lists
