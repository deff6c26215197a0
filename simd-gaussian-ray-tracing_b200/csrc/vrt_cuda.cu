// vrt_cuda.cu -- B200 (sm_100a) render path of the `vrt` Gaussian ray tracer behind include/vrt_cuda.h.
//
// Replaces, from scratch and warp-first rather than SIMD-first:
//   K0  k0_prepare      per-Gaussian frame constants + the tiling projection   (src/vrt/rt.cpp:35-45)
//   K1  k1_cull         per-cell Gaussian lists, warp ballot + popc compaction (src/vrt/rt.cpp:47-66)
//       k1_scan/k1_order exclusive scan of the counts, cost-ordered work queue
//   K2  k2_render       per-pixel radiance, closed-form erf transmittance      (src/vrt/rt.h:102-127, 205-223)
//   K3  (K2 epilogue)   clamp / quantise / pack 0xAARRGGBB / vector stores      (src/vrt/rt.h:329-333, 373-377, 388-399)
//
// Work decomposition: one WARP owns one 8x4-pixel cell (lane = pixel) and its own Gaussian list; warps are
// persistent and pull cells from a cost-ordered atomic queue, so there is no block-level barrier anywhere in
// the render kernel.  See DESIGN.md for the algebra (sample-invariant terms hoisted out of the n^2 loop) and
// the roofline.
//
// No CPU fallback: every entry point fails without a CUDA device.  Nothing under oracle/ is used here.
#include "vrt_cuda.h"
#include "vrt_approx_tables.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------------
namespace
{
constexpr int CELL_W = 8;          // pixels per cell (= one warp), x
constexpr int CELL_H = 4;          // y
constexpr int ROOT_SEG = 4096;     // Gaussians per root segment in the first cull level
constexpr int K2_WARPS = 8;        // warps per render CTA
// CTA shape per variant: Q = 8 with a 3-CTA/SM target uses 4-warp CTAs (register cap 168, 12 warps/SM)
__host__ __device__ constexpr int k2_cta_warps(int q, int minb) { return (q == 8 && minb == 3) ? 4 : K2_WARPS; }
constexpr int STAGE = 32;          // records staged per warp per step (one per lane)
constexpr int WIN_CAP = 160;       // longest list the depth-window kernel caches per warp ((WIN_CAP+1) * 128 B of prefix sums)
// Heavy cells (work ~ n^2) are split by emitter range into independent work items so one warp never owns a whole long list:
// a cell with more than 3 x slice entries becomes ceil(n / slice) items; the partial radiances are summed in slice order.
constexpr int SLICE_MAX = 64;      // emitters per item of a split cell: 64 on big frames, down to 8 when a frame has too few
constexpr int SLICE_MIN = 8;       //   items to fill the machine (always a multiple of every emitter block size Q)
constexpr int ITEM_CELL_BITS = 22; // work item = cell id | slice << 22  (4M cells, 1024 slices)
constexpr uint32_t NO_SLOT = 0xFFFFFFFFu;
// Literal list modes: a Gaussian farther than this many sigma from every ray of a cell has weight exp(-d^2 / 2 sigma^2) <
// 2^-126, which MUFU.EX2 (.ftz) returns as exactly 0 -- it contributes exactly 0 to every sum of the cell (13.22 sigma is the
// exact limit; the margin covers fp32 rounding of d^2 and fast_exp's clamp at 13.27 sigma).
constexpr float VISIBLE_SIGMAS = 13.4f;

constexpr float LOG2E = 1.4426950408889634f;
constexpr float SQRT_PI_2 = 1.2533141373155003f; // sqrt(pi/2) = 1/0.7978845608 (INV_SQRT_2_PI of src/vrt/rt.h:19)
constexpr float REF_CULL_SIGMAS = 3.3f;           // src/vrt/rt.cpp:58-59

struct FrameGeom
{
    // camera
    float inv0[3], inv1[3], inv3[3]; // columns 0, 1, 3 of inverse(view) (xyz)
    float origin[3];
    float view[16];
    // image
    int W, H;
    int tiles_x, tiles_y, tile_w, tile_h;
    int cptx, cpty;   // cells per tile
    int ncx, ncy;     // global cell grid
    int row_begin, row_end;
    int slice;        // emitters per work item of a split cell (build_queue picks it per frame); cells with <= 3 slice entries stay whole
    int uniform;      // every tile is a whole number of cells and cells tile the image exactly: cell (cx, cy) starts at (8 cx, 4 cy)
    // list semantics
    int use_ref;      // apply the reference predicate
    int use_bound;    // apply the per-cell k-sigma bound
    int list_kind;    // 0: per-cell lists (index), 1: per-tile lists, 2: single list (all)
    float bound_k;
    float tw, th;     // 2/tiles
    float half_w, half_h; // W/2, H/2 as float
};

__constant__ FrameGeom c_geom;
__constant__ float c_tile_cx[1024]; // float-accumulated tile centres (src/vrt/rt.cpp:47-49)
__constant__ float c_tile_cy[1024];

// per-Gaussian frame record: 3 x float4
//   a = (oc.x, oc.y, oc.z, (mu.w - o.w)^2)            oc = mu - origin
//   b = (r = 1/(sqrt2 sigma), r2l = log2e/(2 sigma^2), Kl = sigma c sqrt(pi/2) log2e, sigma)
//   c = albedo xyzw
struct alignas(16) Rec
{
    float4 a, b, c;
};

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float copysign_bits(float mag, float sgn)
{
    // (mag & 0x7fffffff) | (sgn & 0x80000000): one LOP3 on the ALU pipe
    return __uint_as_float((__float_as_uint(mag) & 0x7fffffffu) | (__float_as_uint(sgn) & 0x80000000u));
}

// Abramowitz-Stegun 7.1.27 with the reference's coefficients (src/vrt/approx.cpp:90-110):
//   erf(x) = sign(x) (1 - 1/(1 + a1|x| + a2 x^2 + a3 |x|^3 + a4 x^4)^4)
// 4 FFMA + 2 FMUL + MUFU.RCP + FADD on the FMA/XU pipes, |x| and the sign transfer on the ALU pipe.
constexpr float AS_A1 = 0.278393f, AS_A2 = 0.230389f, AS_A3 = 0.000972f, AS_A4 = 0.078108f;

__device__ __forceinline__ float erf_as(float t)
{
    const float x = fabsf(t);
    float d = fmaf(AS_A4, x, AS_A3);
    d = fmaf(d, x, AS_A2);
    d = fmaf(d, x, AS_A1);
    d = fmaf(d, x, 1.f);
    d = d * d;
    d = d * d;
    return copysign_bits(1.f - rcp_approx(d), t);
}

// libm-class erf on the FMA pipe: erf(|x|) = 1 - 2^(-|x| P(|x|)), P of degree 6 fitted on [0, 4]
// (tools/fit_erf.py: max abs error 7.7e-8 in exact arithmetic, 1.7e-7 with fp32 Horner, against double
// erf; |x| is clamped to 4 where the form returns 1 - 1.7e-8).  6 FFMA + FMUL + MUFU.EX2 + FADD.
constexpr float EX_XMAX = 4.0f;
constexpr float EX_C0 = 1.6279137324e+00f, EX_C1 = 9.1832863539e-01f, EX_C2 = 1.4896371499e-01f, EX_C3 = -2.9452616825e-02f,
                EX_C4 = 2.3023453175e-03f, EX_C5 = 4.6152042132e-04f, EX_C6 = -1.0021147713e-04f;

__device__ __forceinline__ float erf_exact(float t)
{
    const float x = fminf(fabsf(t), EX_XMAX);
    float p = fmaf(EX_C6, x, EX_C5);
    p = fmaf(p, x, EX_C4);
    p = fmaf(p, x, EX_C3);
    p = fmaf(p, x, EX_C2);
    p = fmaf(p, x, EX_C1);
    p = fmaf(p, x, EX_C0);
    return copysign_bits(1.f - ex2_approx(-p * x), t);
}

template <int ERF>
__device__ __forceinline__ float erf_variant(float t)
{
    return ERF == 0 ? erf_as(t) : erf_exact(t);
}

// Packed (2 x fp32) forms: Blackwell issues FFMA2 / FMUL2 / FADD2 on 64-bit register pairs, halving the
// issue slots of the FMA-pipe part of the inner term (the loop is issue-bound in scalar form).
template <int ERF>
/// w(t) = 1 - |erf(t)|, the even part both variants compute first: 1/D(|t|)^4 (A&S) or 2^(-|t| P(|t|)) (exact).
__device__ __forceinline__ float2 erfc_mag2(float2 t)
{
    if (ERF == 0)
    {
        const float2 x = make_float2(fabsf(t.x), fabsf(t.y));
        float2 d = __ffma2_rn(make_float2(AS_A4, AS_A4), x, make_float2(AS_A3, AS_A3));
        d = __ffma2_rn(d, x, make_float2(AS_A2, AS_A2));
        d = __ffma2_rn(d, x, make_float2(AS_A1, AS_A1));
        d = __ffma2_rn(d, x, make_float2(1.f, 1.f));
        d = __fmul2_rn(d, d);
        d = __fmul2_rn(d, d);
        return make_float2(rcp_approx(d.x), rcp_approx(d.y));
    }
    else
    {
        const float2 x = make_float2(fminf(fabsf(t.x), EX_XMAX), fminf(fabsf(t.y), EX_XMAX));
        float2 p = __ffma2_rn(make_float2(EX_C6, EX_C6), x, make_float2(EX_C5, EX_C5));
        p = __ffma2_rn(p, x, make_float2(EX_C4, EX_C4));
        p = __ffma2_rn(p, x, make_float2(EX_C3, EX_C3));
        p = __ffma2_rn(p, x, make_float2(EX_C2, EX_C2));
        p = __ffma2_rn(p, x, make_float2(EX_C1, EX_C1));
        p = __ffma2_rn(p, x, make_float2(EX_C0, EX_C0));
        const float2 q = __fmul2_rn(p, make_float2(-x.x, -x.y));
        return make_float2(ex2_approx(q.x), ex2_approx(q.y));
    }
}

template <int ERF>
__device__ __forceinline__ float2 erf_variant2(float2 t)
{
    const float2 w = erfc_mag2<ERF>(t);
    const float2 v = __ffma2_rn(w, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
    return make_float2(copysign_bits(v.x, t.x), copysign_bits(v.y, t.y));
}

// ------------------------------------------------------------------------------------------------
// K0: per-Gaussian frame constants
// ------------------------------------------------------------------------------------------------
// cull record, 32 B: (oc.xyz, sigma) and (mu'.x, mu'.y, 3.3 sigma', valid) of the reference's tiling projection (rt.cpp:35-45)
__global__ void k0_prepare(const float *__restrict__ aos, uint64_t n, Rec *__restrict__ rec, float4 *__restrict__ cullrec)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *g = aos + i * 10;
    const float ax = g[0], ay = g[1], az = g[2], aw = g[3];
    const float mx = g[4], my = g[5], mz = g[6], mw = g[7];
    const float sigma = g[8], mag = g[9];
    Rec r;
    r.a = make_float4(mx - c_geom.origin[0], my - c_geom.origin[1], mz - c_geom.origin[2], mw * mw);
    const float rr = 1.f / (1.41421356237309504880f * sigma);
    r.b = make_float4(rr, LOG2E / (2.f * sigma * sigma), sigma * mag * SQRT_PI_2 * LOG2E, sigma);
    r.c = make_float4(ax, ay, az, aw);
    rec[i] = r;
    if (cullrec != nullptr)
    {
        // proj = view * (mu.xyz, 1), GLM operand order (c0 x + c1 y) + (c2 z + c3 w)
        const float *v = c_geom.view;
        const float px = (v[0] * mx + v[4] * my) + (v[8] * mz + v[12]);
        const float py = (v[1] * mx + v[5] * my) + (v[9] * mz + v[13]);
        const float pz = (v[2] * mx + v[6] * my) + (v[10] * mz + v[14]);
        const float inv = 1.f / pz;
        const float sg = sigma * inv;
        const bool valid = !(pz < 1.f) && !(sg < 1e-5f);
        // one 32-byte sector per Gaussian holds everything K1 tests: (oc.xyz, sigma) and the reference projection
        cullrec[2 * i] = make_float4(r.a.x, r.a.y, r.a.z, sigma);
        cullrec[2 * i + 1] = make_float4(px * inv, py * inv, REF_CULL_SIGMAS * sg, valid ? 1.f : 0.f);
    }
}

// ------------------------------------------------------------------------------------------------
// K1: culling
// ------------------------------------------------------------------------------------------------
struct CullRect
{
    // outward unit normals of the four side planes of the rect's ray frustum (apex = origin)
    float nl[3], nr[3], nb[3], nt[3];
    float u0, u1, v0, v1;   // plane coordinates of the extreme pixel samples (the corner rays)
    int tx0, tx1, ty0, ty1; // reference tile range covered
    bool exact_tile;        // single tile: evaluate the predicate exactly
};

__device__ __forceinline__ void cross3(const float *a, const float *b, float *r)
{
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void orient_normalize(float *n, const float *towards, float sign)
{
    const float inv = rsqrtf(fmaxf(dot3(n, n), 1e-30f));
    const float s = (dot3(n, towards) * sign >= 0.f) ? inv : -inv;
    n[0] *= s; n[1] *= s; n[2] *= s;
}

// pixel rect [x0,x1) x [y0,y1) -> frustum planes through the extreme sample positions
__device__ __forceinline__ void make_rect(int x0, int x1, int y0, int y1, CullRect &rc)
{
    const FrameGeom &G = c_geom;
    rc.u0 = -1.f + (float)x0 / G.half_w; rc.u1 = -1.f + (float)(x1 - 1) / G.half_w;
    rc.v0 = -1.f + (float)y0 / G.half_h; rc.v1 = -1.f + (float)(y1 - 1) / G.half_h;
    float Wv[3], a[3];
    for (int i = 0; i < 3; ++i) Wv[i] = G.inv3[i] - G.origin[i];
    // left / right planes contain U = inv1 and the ray (u * inv0 + Wv)
    for (int i = 0; i < 3; ++i) a[i] = rc.u0 * G.inv0[i] + Wv[i];
    cross3(G.inv1, a, rc.nl); orient_normalize(rc.nl, G.inv0, -1.f);
    for (int i = 0; i < 3; ++i) a[i] = rc.u1 * G.inv0[i] + Wv[i];
    cross3(G.inv1, a, rc.nr); orient_normalize(rc.nr, G.inv0, +1.f);
    // bottom / top planes contain R = inv0 and the ray (v * inv1 + Wv)
    for (int i = 0; i < 3; ++i) a[i] = rc.v0 * G.inv1[i] + Wv[i];
    cross3(G.inv0, a, rc.nb); orient_normalize(rc.nb, G.inv1, -1.f);
    for (int i = 0; i < 3; ++i) a[i] = rc.v1 * G.inv1[i] + Wv[i];
    cross3(G.inv0, a, rc.nt); orient_normalize(rc.nt, G.inv1, +1.f);
    rc.tx0 = rc.tx1 = rc.ty0 = rc.ty1 = 0;
    if (G.use_ref)
    {
        rc.tx0 = x0 / G.tile_w; rc.tx1 = (x1 - 1) / G.tile_w;
        rc.ty0 = y0 / G.tile_h; rc.ty1 = (y1 - 1) / G.tile_h;
    }
    rc.exact_tile = (rc.tx0 == rc.tx1) && (rc.ty0 == rc.ty1);
}

// reference predicate for one axis (src/vrt/rt.cpp:57-59): |c - mu'| <= |c| + t/2 + 3.3 sigma'
__device__ __forceinline__ bool ref_axis(float c, float mu, float half_t, float s33) { return fabsf(c - mu) <= fabsf(c) + half_t + s33; }

// Distance test against one orientation of the frustum (sgn = +1: the frustum itself, -1: its mirror image through the
// apex).  Inside the k-sigma slab of all four planes; a centre outside TWO adjacent planes (distances su, sv > 0, cosine c
// between their normals) is nearest to the corner ray only if it projects beyond the edge on BOTH faces (su - sv c > 0 and
// sv - su c > 0) -- then its distance to that ray's line decides (rounded corners instead of a box: ~14 % shorter lists);
// otherwise a face is nearest and its plane distance (already <= lim) is the true distance.
__device__ __forceinline__ bool near_frustum(const CullRect &rc, const float *p, float dl, float dr, float db, float dt, float sgn, float lim)
{
    const FrameGeom &G = c_geom;
    const float sl = sgn * dl, sr = sgn * dr, sb = sgn * db, st = sgn * dt;
    const float su = fmaxf(sl, sr), sv = fmaxf(sb, st);
    if (!(su <= lim && sv <= lim)) return false;
    if (!(su > 0.f && sv > 0.f)) return true;
    const bool right = sr > sl, top = st > sb;
    float c = 0.f;
    for (int i = 0; i < 3; ++i) c += (right ? rc.nr[i] : rc.nl[i]) * (top ? rc.nt[i] : rc.nb[i]);
    if (!(su - sv * c > 0.f && sv - su * c > 0.f)) return true;
    const float uu = right ? rc.u1 : rc.u0, vv = top ? rc.v1 : rc.v0;
    float e[3];
    for (int i = 0; i < 3; ++i) e[i] = uu * G.inv0[i] + vv * G.inv1[i] + (G.inv3[i] - G.origin[i]);
    const float t = __fdividef(dot3(p, e), dot3(e, e));
    const float px = p[0] - t * e[0], py = p[1] - t * e[1], pz = p[2] - t * e[2];
    return px * px + py * py + pz * pz <= lim * lim;
}

__device__ __forceinline__ bool cull_test(const CullRect &rc, const float4 a, const float sigma, const float4 cr)
{
    const FrameGeom &G = c_geom;
    if (G.use_ref)
    {
        if (cr.w == 0.f) return false;
        const float hx = G.tw / 2, hy = G.th / 2;
        if (rc.exact_tile)
        {
            if (!(ref_axis(c_tile_cx[rc.tx0], cr.x, hx, cr.z) && ref_axis(c_tile_cy[rc.ty0], cr.y, hy, cr.z))) return false;
        }
        else
        {
            // |c - mu| - |c| is monotone in c, so a tile range passes iff one of its end tiles does;
            // the slack keeps the coarse level conservative against rounding of the exact test.
            const float s = cr.z + 1e-4f;
            const bool px = ref_axis(c_tile_cx[rc.tx0], cr.x, hx, s) || ref_axis(c_tile_cx[rc.tx1], cr.x, hx, s);
            const bool py = ref_axis(c_tile_cy[rc.ty0], cr.y, hy, s) || ref_axis(c_tile_cy[rc.ty1], cr.y, hy, s);
            if (!(px && py)) return false;
        }
    }
    if (G.use_bound)
    {
        const float p[3] = {a.x, a.y, a.z};
        // distance budget: k sigma, plus the w offset can only increase the true distance (ignored => conservative)
        const float lim = G.bound_k * sigma + 1e-6f * (fabsf(a.x) + fabsf(a.y) + fabsf(a.z));
        const float dl = dot3(p, rc.nl), dr = dot3(p, rc.nr), db = dot3(p, rc.nb), dt = dot3(p, rc.nt);
        // the reference integrates along the whole line (samples with s < 0 are not guarded, rt.h:155-160),
        // so the mirrored frustum counts too
        if (!(near_frustum(rc, p, dl, dr, db, dt, 1.f, lim) || near_frustum(rc, p, dl, dr, db, dt, -1.f, lim))) return false;
    }
    return true;
}

// One culling level.  The cell grid is grouped into gx x gy-cell groups (ngx x ngy of them); every group scans the list of
// the coarser group that contains it (pgx x pgy cells, pngx per row) and keeps what passes its own test.  The root level
// has no parent: it scans the scene itself in n_seg segments of ROOT_SEG Gaussians, one warp per (group, segment), and the
// per-segment pieces concatenate in index order.  The finest level has gx = gy = 1 (one 8x4-pixel cell per warp).
struct CullLevel
{
    int gx, gy, ngx, ngy;
    int pgx, pgy, pngx;
    int is_root, n_seg;
};

// WRITE = false: counts[group * n_seg + seg] ; WRITE = true: indices at offsets[group * n_seg + seg].
// 32 candidates per step, one per lane: predicate -> ballot -> popc of the lower lanes = ordered slot (lists keep
// ascending Gaussian index, so K2's sums are reproducible).
template <bool WRITE>
__global__ void __launch_bounds__(256) k1_cull(const Rec *__restrict__ rec, const float4 *__restrict__ cullrec, uint32_t n_root, const CullLevel L,
                                               const uint32_t *__restrict__ parent_off, const uint32_t *__restrict__ parent_idx,
                                               uint32_t *__restrict__ counts, const uint32_t *__restrict__ offsets, uint32_t *__restrict__ out_idx,
                                               uint32_t n_work)
{
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= n_work) return;
    const FrameGeom &G = c_geom;
    const uint32_t group = wid / L.n_seg, seg = wid % L.n_seg;
    const int gxi = group % L.ngx, gyi = group / L.ngx;
    // pixel rect of the group = union of its cells' rects
    const int cx0 = gxi * L.gx, cx1 = min(G.ncx, cx0 + L.gx) - 1;
    const int cy0 = gyi * L.gy, cy1 = min(G.ncy, cy0 + L.gy) - 1;
    int x0, y0, x1, y1;
    if (G.uniform)
    {
        x0 = cx0 * CELL_W; y0 = cy0 * CELL_H;
        x1 = (cx1 + 1) * CELL_W; y1 = (cy1 + 1) * CELL_H;
    }
    else
    {
        // ragged tiles (tile size not a multiple of the cell): cells restart at every tile edge
        x0 = (cx0 / G.cptx) * G.tile_w + (cx0 % G.cptx) * CELL_W;
        y0 = (cy0 / G.cpty) * G.tile_h + (cy0 % G.cpty) * CELL_H;
        x1 = min((cx1 / G.cptx) * G.tile_w + min(G.tile_w, (cx1 % G.cptx + 1) * CELL_W), G.W);
        y1 = min((cy1 / G.cpty) * G.tile_h + min(G.tile_h, (cy1 % G.cpty + 1) * CELL_H), G.H);
    }
    uint32_t begin, end;
    if (L.is_root)
    {
        begin = seg * ROOT_SEG;
        end = min(n_root, begin + ROOT_SEG);
    }
    else
    {
        const uint32_t parent = (uint32_t)((cy0 / L.pgy) * L.pngx + (cx0 / L.pgx));
        begin = parent_off[parent];
        end = parent_off[parent + 1];
    }
    // groups outside the rendered row band get empty lists
    const bool in_band = y1 > G.row_begin && y0 < G.row_end;
    uint32_t base = WRITE ? offsets[wid] : 0u;
    uint32_t count = 0;
    if (in_band && end > begin)
    {
        CullRect rc;
        make_rect(x0, x1, y0, y1, rc);
        for (uint32_t k = begin; k < end; k += 32)
        {
            const uint32_t e = k + lane;
            bool pass = false;
            uint32_t gi = 0;
            if (e < end)
            {
                gi = (L.is_root || parent_idx == nullptr) ? e : parent_idx[e]; // no index array: the parent list is a contiguous range
                const float4 a = cullrec[2 * gi]; // (oc.xyz, sigma)
                const float4 cr = G.use_ref ? cullrec[2 * gi + 1] : make_float4(0.f, 0.f, 0.f, 1.f);
                pass = cull_test(rc, a, a.w, cr);
            }
            const uint32_t ballot = __ballot_sync(0xffffffffu, pass);
            if (WRITE)
            {
                if (pass) out_idx[base + __popc(ballot & ((1u << lane) - 1u))] = gi;
                base += __popc(ballot);
            }
            else count += __popc(ballot);
        }
    }
    if (!WRITE && lane == 0) counts[wid] = count;
}

// offsets of a segmented root level -> one offset per group (+ the total)
__global__ void k1_group_offsets(const uint32_t *__restrict__ seg_offsets, uint32_t *__restrict__ group_offsets, uint32_t n_groups, int n_seg)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_groups) group_offsets[i] = seg_offsets[(size_t)i * n_seg];
}

// pure REFERENCE lists (one list per reference tile), level 1 with children = tiles
template <bool WRITE>
__global__ void __launch_bounds__(256) k1_cull_tiles(const Rec *__restrict__ rec, const float4 *__restrict__ cullrec, uint32_t n_root,
                                                     uint32_t *__restrict__ counts, const uint32_t *__restrict__ offsets,
                                                     uint32_t *__restrict__ out_idx, uint32_t n_tiles)
{
    // one CTA per tile; ordered compaction across the CTA's 8 warps
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    const FrameGeom &G = c_geom;
    const uint32_t tile = blockIdx.x;
    if (tile >= n_tiles) return;
    const int tx = tile % G.tiles_x, ty = tile / G.tiles_x;
    const float cx = c_tile_cx[tx], cy = c_tile_cy[ty];
    const float hx = G.tw / 2, hy = G.th / 2;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = WRITE ? offsets[tile] : 0u;
    __syncthreads();
    for (uint32_t k = 0; k < n_root; k += 256)
    {
        const uint32_t i = k + threadIdx.x;
        bool pass = false;
        if (i < n_root)
        {
            const float4 cr = cullrec[2 * i + 1];
            pass = cr.w != 0.f && ref_axis(cx, cr.x, hx, cr.z) && ref_axis(cy, cr.y, hy, cr.z);
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, pass);
        if (lane == 0) s_warp[w] = __popc(ballot);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int q = 0; q < 8; ++q)
        {
            const uint32_t c = s_warp[q];
            if (q < w) before += c;
            total += c;
        }
        if (WRITE && pass) out_idx[s_base + before + __popc(ballot & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
    }
    __syncthreads();
    if (!WRITE && threadIdx.x == 0) counts[tile] = s_base;
}

// Depth order for the depth-window mode: every cell's index list is sorted by the depth of the centre along the cell's
// centre ray (ties by Gaussian index, so the order is deterministic).  One warp per cell, bitonic network in shared memory;
// lists longer than SORT_CAP stay in index order (the window test is valid for any order, it just saturates less often).
constexpr int SORT_CAP = 512;
__device__ __forceinline__ void cell_rect(int cx, int cy, int &x0, int &y0, int &w, int &h);
__global__ void __launch_bounds__(128) k1_sort_cells(const float4 *__restrict__ cullrec, const uint32_t *__restrict__ list_off, uint32_t *__restrict__ list_idx,
                                                     uint32_t n_cells)
{
    __shared__ float s_key[4][SORT_CAP];
    __shared__ uint32_t s_val[4][SORT_CAP];
    const FrameGeom &G = c_geom;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t cell = blockIdx.x * 4 + w;
    if (cell >= n_cells) return;
    const uint32_t off = list_off[cell], n = list_off[cell + 1] - off;
    if (n < 2 || n > SORT_CAP) return;
    int x0, y0, cw, ch;
    cell_rect((int)(cell % G.ncx), (int)(cell / G.ncx), x0, y0, cw, ch);
    // centre ray of the cell
    const float u = -1.f + ((float)x0 + 0.5f * (float)(cw - 1)) / G.half_w, v = -1.f + ((float)y0 + 0.5f * (float)(ch - 1)) / G.half_h;
    float d[3];
    for (int i = 0; i < 3; ++i) d[i] = (G.inv0[i] * u + G.inv1[i] * v) + G.inv3[i] - G.origin[i];
    const float inv = rsqrtf(fmaxf(dot3(d, d), 1e-30f));
    uint32_t m = 2;
    while (m < n) m <<= 1;
    float *key = s_key[w];
    uint32_t *val = s_val[w];
    for (uint32_t i = lane; i < m; i += 32)
    {
        if (i < n)
        {
            const uint32_t gi = list_idx[off + i];
            const float4 a = cullrec[2 * gi];
            key[i] = (a.x * d[0] + a.y * d[1] + a.z * d[2]) * inv;
            val[i] = gi;
        }
        else
        {
            key[i] = 3.0e38f;
            val[i] = 0xFFFFFFFFu;
        }
    }
    __syncwarp();
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1)
        {
            for (uint32_t i = lane; i < m; i += 32)
            {
                const uint32_t l = i ^ j;
                if (l > i)
                {
                    const float ki = key[i], kl = key[l];
                    const uint32_t vi = val[i], vl = val[l];
                    const bool up = (i & k) == 0;
                    const bool gt = ki > kl || (ki == kl && vi > vl);
                    if (gt == up)
                    {
                        key[i] = kl; key[l] = ki;
                        val[i] = vl; val[l] = vi;
                    }
                }
            }
            __syncwarp();
        }
    for (uint32_t i = lane; i < n; i += 32) list_idx[off + i] = val[i];
}

// exclusive scan of n counts into n+1 offsets; single CTA of 1024 threads (n <= a few million)
__global__ void __launch_bounds__(1024) k1_scan(const uint32_t *__restrict__ counts, uint32_t *__restrict__ offsets, uint32_t n)
{
    __shared__ uint32_t s_part[1024];
    const uint32_t per = (n + 1023u) / 1024u;
    const uint32_t b = threadIdx.x * per, e = min(n, b + per);
    uint32_t sum = 0;
    for (uint32_t i = b; i < e; ++i) sum += counts[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 partials
    for (int d = 1; d < 1024; d <<= 1)
    {
        const uint32_t v = (threadIdx.x >= (uint32_t)d) ? s_part[threadIdx.x - d] : 0u;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = s_part[threadIdx.x] - sum;
    for (uint32_t i = b; i < e; ++i)
    {
        offsets[i] = run;
        run += counts[i];
    }
    if (threadIdx.x == 1023) offsets[n] = s_part[1023];
}

// Large arrays are scanned in three launches: per-tile sums (SCAN_TILE elements per CTA), k1_scan over the tile sums,
// then every CTA rescans its tile starting from its tile offset.
constexpr int SCAN_TILE = 4096; // 256 threads x 16 consecutive elements
__device__ __forceinline__ uint32_t block_exclusive_256(uint32_t v, uint32_t *s_warp, uint32_t &block_total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (int q = 0; q < 8; ++q)
    {
        const uint32_t c = s_warp[q];
        if (q < w) before += c;
        total += c;
    }
    block_total = total;
    return before + inc - v;
}

__global__ void __launch_bounds__(256) k1_scan_tiles(const uint32_t *__restrict__ counts, uint32_t *__restrict__ tile_sums, uint32_t n)
{
    __shared__ uint32_t s_warp[8];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * 16;
    uint32_t sum = 0;
    for (int i = 0; i < 16; ++i)
        if (base + i < n) sum += counts[base + i];
    uint32_t total;
    block_exclusive_256(sum, s_warp, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) k1_scan_apply(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ tile_offsets,
                                                     uint32_t *__restrict__ offsets, uint32_t n, uint32_t n_tiles)
{
    __shared__ uint32_t s_warp[8];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * 16;
    uint32_t v[16], sum = 0;
    for (int i = 0; i < 16; ++i)
    {
        v[i] = base + i < n ? counts[base + i] : 0u;
        sum += v[i];
    }
    uint32_t total;
    uint32_t run = tile_offsets[blockIdx.x] + block_exclusive_256(sum, s_warp, total);
    for (int i = 0; i < 16; ++i)
    {
        if (base + i < n) offsets[base + i] = run;
        run += v[i];
    }
    if (blockIdx.x == n_tiles - 1 && threadIdx.x == 255) offsets[n] = tile_offsets[n_tiles]; // grand total
}

// statistics + cost histogram of the render cells.  list id of a cell: per-cell lists -> cell, per-tile ->
// its tile, single -> 0.  key = min(n, 65535); the queue is filled in descending key order.
struct TileStats
{
    unsigned long long entries;   // sum n over lists
    unsigned long long max_list;
    double terms_listed;          // sum over band pixels of 5 n^2
    unsigned long long terms_exec; // filled by K2
    unsigned long long terms_sat;  // K2, depth-window mode: terms resolved by saturation
    unsigned long long n_big;      // queued items whose list is longer than the depth-window cache
    unsigned long long n_items;    // work items queued (cells + extra slices of split cells)
    unsigned long long n_split;    // items that belong to split cells (= partial-radiance slots)
};

__device__ __forceinline__ uint32_t cell_list_id(int cx, int cy)
{
    const FrameGeom &G = c_geom;
    if (G.list_kind == 0) return (uint32_t)(cy * G.ncx + cx);
    if (G.list_kind == 1) return (uint32_t)((cy / G.cpty) * G.tiles_x + (cx / G.cptx));
    return 0u;
}

__device__ __forceinline__ void cell_rect(int cx, int cy, int &x0, int &y0, int &w, int &h)
{
    const FrameGeom &G = c_geom;
    if (G.uniform)
    {
        x0 = cx * CELL_W; y0 = cy * CELL_H;
        w = CELL_W; h = CELL_H;
        return;
    }
    const int lx = (cx % G.cptx) * CELL_W, ly = (cy % G.cpty) * CELL_H;
    x0 = (cx / G.cptx) * G.tile_w + lx;
    y0 = (cy / G.cpty) * G.tile_h + ly;
    w = min(CELL_W, G.tile_w - lx);
    h = min(CELL_H, G.tile_h - ly);
}

// number of work items of a cell with an n-entry list
__device__ __forceinline__ uint32_t cell_items(uint32_t n, uint32_t cell)
{
    const uint32_t slice = (uint32_t)c_geom.slice;
    if (n <= 3u * slice || cell >= (1u << ITEM_CELL_BITS)) return 1u;
    const uint32_t k = (n + slice - 1) / slice;
    return k <= (1u << (32 - ITEM_CELL_BITS)) ? k : 1u;
}

// COUNT_ITEMS = false: listed terms and per-row cost; true: work items per list-length key (needs c_geom.slice)
template <bool COUNT_ITEMS>
__global__ void k1_hist(const uint32_t *__restrict__ list_off, uint32_t *__restrict__ hist, TileStats *__restrict__ stats,
                        double *__restrict__ row_cost, int cy_begin, int cy_end)
{
    const FrameGeom &G = c_geom;
    const int ncells = (cy_end - cy_begin) * G.ncx;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double terms = 0.0;
    if (i < ncells)
    {
        const int cx = i % G.ncx, cy = cy_begin + i / G.ncx;
        int x0, y0, w, h;
        cell_rect(cx, cy, x0, y0, w, h);
        const int ya = max(y0, G.row_begin), yb = min(y0 + h, G.row_end);
        const uint32_t id = cell_list_id(cx, cy);
        const uint32_t n = list_off[id + 1] - list_off[id];
        if (yb > ya)
        {
            if (COUNT_ITEMS)
            {
                const uint32_t items = cell_items(n, (uint32_t)(cy * G.ncx + cx));
                atomicAdd(&hist[min(n, 65535u)], items);
                atomicAdd(&stats->n_items, (unsigned long long)items);
                if (items > 1) atomicAdd(&stats->n_split, (unsigned long long)items);
            }
            else
            {
                terms = 5.0 * (double)n * (double)n * (double)(w * (yb - ya));
                if (row_cost != nullptr) atomicAdd(&row_cost[cy], terms);
            }
        }
    }
    if (!COUNT_ITEMS)
    {
        // warp reduction of the listed terms
        for (int o = 16; o > 0; o >>= 1) terms += __shfl_xor_sync(0xffffffffu, terms, o);
        if ((threadIdx.x & 31) == 0 && terms != 0.0) atomicAdd(&stats->terms_listed, terms);
    }
}

__global__ void k1_list_stats(const uint32_t *__restrict__ list_off, uint32_t n_lists, TileStats *__restrict__ stats)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = 0;
    if (i < n_lists) n = list_off[i + 1] - list_off[i];
    uint32_t mx = n;
    unsigned long long sum = n;
    for (int o = 16; o > 0; o >>= 1)
    {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if ((threadIdx.x & 31) == 0 && sum)
    {
        atomicAdd(&stats->entries, sum);
        atomicMax(&stats->max_list, (unsigned long long)mx);
    }
}

// hist (ascending key) -> start position of each key in a DESCENDING ordering; single CTA
__global__ void __launch_bounds__(1024) k1_hist_scan(uint32_t *__restrict__ hist)
{
    __shared__ uint32_t s_part[1024];
    // thread t owns keys [t*64, t*64+64) ; descending order => process from the top
    const int t = threadIdx.x;
    uint32_t sum = 0;
    for (int k = 0; k < 64; ++k) sum += hist[65535 - (t * 64 + k)];
    s_part[t] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1)
    {
        const uint32_t v = (t >= d) ? s_part[t - d] : 0u;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    uint32_t run = s_part[t] - sum;
    for (int k = 0; k < 64; ++k)
    {
        const int key = 65535 - (t * 64 + k);
        const uint32_t c = hist[key];
        hist[key] = run;
        run += c;
    }
}

__global__ void k1_order(const uint32_t *__restrict__ list_off, uint32_t *__restrict__ cursor, uint32_t *__restrict__ queue, uint32_t *__restrict__ cell_slot,
                         uint32_t *__restrict__ split_cursor, int cy_begin, int cy_end)
{
    const FrameGeom &G = c_geom;
    const int ncells = (cy_end - cy_begin) * G.ncx;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncells) return;
    const int cx = i % G.ncx, cy = cy_begin + i / G.ncx;
    int x0, y0, w, h;
    cell_rect(cx, cy, x0, y0, w, h);
    if (min(y0 + h, G.row_end) <= max(y0, G.row_begin)) return;
    const uint32_t id = cell_list_id(cx, cy);
    const uint32_t n = list_off[id + 1] - list_off[id];
    const uint32_t cell = (uint32_t)(cy * G.ncx + cx);
    const uint32_t items = cell_items(n, cell);
    const uint32_t pos = atomicAdd(&cursor[min(n, 65535u)], items);
    for (uint32_t k = 0; k < items; ++k) queue[pos + k] = cell | (k << ITEM_CELL_BITS);
    // split cells get `items` consecutive slots of the partial-radiance buffer (the slot order is irrelevant: K3' sums a
    // cell's slices in slice order)
    cell_slot[cell] = items > 1 ? atomicAdd(split_cursor, items) : NO_SLOT;
}

// ------------------------------------------------------------------------------------------------
// K2 + K3: render
// ------------------------------------------------------------------------------------------------
struct RenderArgs
{
    const Rec *rec;            // frame records
    const uint32_t *list_off;  // per list: [off, off+n)
    const uint32_t *list_idx;  // indices into rec, or nullptr when lists are contiguous ranges of rec
    const uint32_t *queue;     // cost-ordered cell ids
    uint32_t n_queue;
    uint32_t *counter;         // work-queue head
    uint32_t *image;           // W*H packed pixels (may be null)
    float4 *radiance;          // W*H float4 (may be null)
    unsigned long long *terms_exec;
    unsigned long long *terms_sat;
    float skip_thresh;         // skip an occluder / emitter for the whole warp when exp2 weight <= thresh (-1: never)
    uint32_t quant_nearest, alpha_from_w;
    uint32_t window; // depth-window mode
    const uint32_t *cell_slot; // per cell: first slot of its slices in `partial`, NO_SLOT for whole cells (may be null)
    float4 *partial;           // [slot][lane] partial radiance of the items of split cells
};

// ---- per-warp record staging -------------------------------------------------------------------------------------------
// Each warp owns two STAGE-record buffers and two mbarriers.  When a list is a contiguous range of `rec` (ALL lists,
// caller-supplied tiles_t lists) a chunk is one TMA bulk copy (cp.async.bulk global -> shared, completion on the mbarrier)
// issued by lane 0 one chunk ahead of the compute; index lists are gathered by the lanes (one record per lane).
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    // try_wait suspends the thread for a hardware time slice per probe; the probe count is bounded so that a protocol
    // error traps instead of hanging the GPU
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 24); ++spin)
    {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    // generic-proxy reads of the buffer (previous chunk) are ordered before the async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

struct PixelRay
{
    float nx, ny, nz;
};

// occluder quantities for this lane's ray: weight A (log2 units), mu_bar, and exp2 factor e
__device__ __forceinline__ void occluder_setup(const float4 a, const float4 b, const PixelRay &ray, float &mu, float &e)
{
    mu = fmaf(a.z, ray.nz, fmaf(a.y, ray.ny, a.x * ray.nx));
    // squared distance from the centre to the ray, from the perpendicular component (no |oc|^2 - mu^2 cancellation)
    const float px = fmaf(-mu, ray.nx, a.x), py = fmaf(-mu, ray.ny, a.y), pz = fmaf(-mu, ray.nz, a.z);
    const float d2 = fmaf(pz, pz, fmaf(py, py, fmaf(px, px, a.w)));
    e = ex2_approx(-d2 * b.y);
}

// K3: clamp, quantise (truncate | round-to-nearest-even) and pack one pixel; store the packed word and/or the float4 radiance
__device__ __forceinline__ void store_pixel(const RenderArgs &args, size_t pi, float Lr, float Lg, float Lb, float La)
{
    if (args.radiance) args.radiance[pi] = make_float4(Lr, Lg, Lb, La);
    if (args.image)
    {
        const float r255 = fminf(Lr, 1.f) * 255.f, g255 = fminf(Lg, 1.f) * 255.f, b255 = fminf(Lb, 1.f) * 255.f;
        uint32_t R, Gc, B, A = 0xFFu;
        if (args.quant_nearest)
        {
            R = (uint32_t)__float2int_rn(r255); Gc = (uint32_t)__float2int_rn(g255); B = (uint32_t)__float2int_rn(b255);
            if (args.alpha_from_w) A = (uint32_t)__float2int_rn(fminf(La, 1.f) * 255.f);
        }
        else
        {
            R = (uint32_t)r255; Gc = (uint32_t)g255; B = (uint32_t)b255;
            if (args.alpha_from_w) A = (uint32_t)(fminf(La, 1.f) * 255.f);
        }
        args.image[pi] = (A << 24) | (R << 16) | (Gc << 8) | B;
    }
}

// ray through pixel (px, py): plane = inverse(view) (u, v, 0, 1) (src/vrt/camera.cpp:60-70); dir = normalize(plane - origin)
__device__ __forceinline__ PixelRay pixel_ray(int px, int py)
{
    const FrameGeom &G = c_geom;
    const float u = -1.f + (float)px / G.half_w, v = -1.f + (float)py / G.half_h;
    const float dx = (G.inv0[0] * u + G.inv1[0] * v) + G.inv3[0] - G.origin[0];
    const float dy = (G.inv0[1] * u + G.inv1[1] * v) + G.inv3[1] - G.origin[1];
    const float dz = (G.inv0[2] * u + G.inv1[2] * v) + G.inv3[2] - G.origin[2];
    const float inv = rsqrtf(dx * dx + dy * dy + dz * dz);
    PixelRay ray;
    ray.nx = dx * inv; ray.ny = dy * inv; ray.nz = dz * inv;
    return ray;
}

template <int ERF, int Q, bool PACK, int MINB, bool CONTIG, bool WIN>
__global__ void __launch_bounds__(k2_cta_warps(Q, MINB) * 32, MINB) k2_render(const RenderArgs args)
{
    constexpr int CTA_WARPS = k2_cta_warps(Q, MINB);
    __shared__ __align__(128) Rec s_rec[CTA_WARPS][2][STAGE];
    __shared__ __align__(8) unsigned long long s_bar[CTA_WARPS][2];
    const FrameGeom &G = c_geom;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = lane & (CELL_W - 1), ly = lane >> 3;
    unsigned long long exec = 0, sat = 0;
    // depth window (WIN): erf saturates to +-esat beyond |t| >= tsat for both variants (A&S: 1 - 1/D^4 rounds to 1.0f from
    // 5.45 on; the exact variant clamps |t| at 4), so an occluder that is that far in front of (behind) EVERY sample of the
    // emitter block for EVERY lane contributes +A esat (-A esat) to all 5Q accumulators: one add instead of 5Q terms
    const float tsat = ERF == 0 ? 5.5f : EX_XMAX;
    const float esat = erf_variant<ERF>(tsat);
    uint32_t par = 0u; // phase parity of this warp's two mbarriers (bit b = buffer b)
    if (CONTIG && lane == 0)
    {
        mbar_init(smem_u32(&s_bar[warp][0]), 1);
        mbar_init(smem_u32(&s_bar[warp][1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    constexpr bool contiguous = CONTIG; // lists are contiguous ranges of rec (TMA) or index lists (gathered by the lanes)

    for (;;)
    {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(args.counter, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= args.n_queue) break;
        const uint32_t item = args.queue[qi];
        const uint32_t cell = item & ((1u << ITEM_CELL_BITS) - 1u), slice = item >> ITEM_CELL_BITS;
        const int cx = cell % G.ncx, cy = cell / G.ncx;
        int x0, y0, cw, ch;
        cell_rect(cx, cy, x0, y0, cw, ch);
        const int px = x0 + min(lx, cw - 1), py = y0 + min(ly, ch - 1);
        const bool live = lx < cw && ly < ch && py >= G.row_begin && py < G.row_end;
        const uint32_t n_live = __popc(__ballot_sync(0xffffffffu, live));

        const PixelRay ray = pixel_ray(px, py);

        const uint32_t lid = cell_list_id(cx, cy);
        const uint32_t off = args.list_off[lid];
        const uint32_t n = args.list_off[lid + 1] - off;

        auto load_rec = [&](uint32_t k) -> const Rec * {
            const uint32_t gi = args.list_idx ? args.list_idx[off + k] : off + k;
            return args.rec + gi;
        };

        // ---- pass A: C = sum_j A_j erf(-m_j)   (the sample-independent half of every term) ----
        const uint32_t n_chunks = (n + STAGE - 1) / STAGE;
        const bool resident = n <= STAGE; // a single chunk stays staged for the whole cell
        auto chunk_count = [&](uint32_t c) { return min((uint32_t)STAGE, n - c * STAGE); };
        // TMA: one bulk copy of the chunk's contiguous records into buffer c & 1, completion on that buffer's mbarrier
        auto issue = [&](uint32_t c) {
            if (lane == 0)
                tma_bulk_load(smem_u32(&s_rec[warp][c & 1][0]), args.rec + off + c * STAGE, chunk_count(c) * (uint32_t)sizeof(Rec), smem_u32(&s_bar[warp][c & 1]));
        };
        // chunk c ready in its buffer; the next chunk is put in flight first
        auto acquire = [&](uint32_t c) -> const Rec * {
            const uint32_t b = c & 1u;
            if (c + 1 < n_chunks) issue(c + 1);
            mbar_wait(smem_u32(&s_bar[warp][b]), (par >> b) & 1u);
            par ^= 1u << b;
            return &s_rec[warp][b][0];
        };
        // index lists: the lanes gather one record each (occluder part only) into buffer 0
        auto gather = [&](uint32_t c) -> const Rec * {
            __syncwarp(); // every lane is done with the previous chunk
            if ((uint32_t)lane < chunk_count(c))
            {
                const Rec *r = args.rec + args.list_idx[off + c * STAGE + lane];
                s_rec[warp][0][lane].a = r->a;
                s_rec[warp][0][lane].b = r->b;
            }
            __syncwarp();
            return &s_rec[warp][0][0];
        };
        auto begin_pass = [&]() {
            if (contiguous && n) issue(0);
        };
        auto chunk_begin = [&](uint32_t c) -> const Rec * { return contiguous ? acquire(c) : gather(c); };
        auto chunk_end = [&](uint32_t c) {
            (void)c;
            if (contiguous) __syncwarp(); // every lane is done with the buffer before TMA refills it
        };
        float C = 0.f;
        const Rec *sr = &s_rec[warp][0][0];
        begin_pass();
        for (uint32_t c = 0; c < n_chunks; ++c)
        {
            sr = chunk_begin(c);
            const uint32_t cnt = chunk_count(c);
            for (uint32_t j = 0; j < cnt; ++j)
            {
                const float4 a = sr[j].a, b = sr[j].b;
                float mu, e;
                occluder_setup(a, b, ray, mu, e);
                C = fmaf(b.z * e, erf_variant<ERF>(-mu * b.x), C);
            }
            if (!resident) chunk_end(c);
        }

        // ---- pass B: emitters in blocks of Q, all occluders per block ----
        float Lr = 0.f, Lg = 0.f, Lb = 0.f, La = 0.f;
        // a split cell's item covers the emitters [q_begin, q_end) only; every item still needs all n occluders
        const uint32_t slot = args.cell_slot ? args.cell_slot[cell] : NO_SLOT;
        const uint32_t q_begin = slot != NO_SLOT ? slice * (uint32_t)G.slice : 0u, q_end = slot != NO_SLOT ? min(n, q_begin + (uint32_t)G.slice) : n;
        for (uint32_t q0 = q_begin; q0 < q_end; q0 += Q)
        {
            // emitter block
            float s[Q][5], acc[Q][5], wgt[Q];
            float4 alb[Q];
            float s0 = 0.f, smin = 3.0e38f, smax = -3.0e38f, base = 0.f;
            bool any_emit = false;
#pragma unroll
            for (int e = 0; e < Q; ++e)
            {
                const bool real = q0 + e < q_end;
                const Rec *r = load_rec(real ? q0 + e : q0);
                const float4 a = r->a, b = r->b;
                alb[e] = r->c;
                float mu, ee;
                occluder_setup(a, b, ray, mu, ee);
                if (e == 0)
                {
                    // one depth shift per warp keeps s r - m small; a degenerate lane-0 ray must not poison the warp
                    s0 = __shfl_sync(0xffffffffu, mu, 0);
                    s0 = (fabsf(s0) <= 3.0e38f) ? s0 : 0.f;
                }
                // emission weight sigma c_bar = Kl e / (sqrt(pi/2) log2e)
                wgt[e] = real ? b.z * ee * (1.f / (SQRT_PI_2 * LOG2E)) : 0.f;
                any_emit |= real && (ee > args.skip_thresh);
                if (real)
                {
                    smin = fminf(smin, (mu - s0) - 4.f * b.w);
                    smax = fmaxf(smax, mu - s0);
                }
#pragma unroll
                for (int k = 0; k < 5; ++k)
                {
                    s[e][k] = (mu - s0) + (float)(k - 4) * b.w;
                    acc[e][k] = 0.f;
                }
            }
            if (!__any_sync(0xffffffffu, any_emit)) continue; // no lane sees any of these emitters
            const uint32_t n_real = min((uint32_t)Q, q_end - q0);

            if (!resident) begin_pass();
            for (uint32_t c = 0; c < n_chunks; ++c)
            {
                const uint32_t cnt = chunk_count(c);
                if (!resident) sr = chunk_begin(c); // lists that fit one chunk stay resident from pass A
                for (uint32_t j = 0; j < cnt; ++j)
                {
                    const float4 a = sr[j].a, b = sr[j].b;
                    float mu, e;
                    occluder_setup(a, b, ray, mu, e);
                    if (!__any_sync(0xffffffffu, e > args.skip_thresh)) continue; // warp-uniform skip
                    const float A = b.z * e;
                    const float r = b.x;
                    const float nm = -(mu - s0) * r;
                    // t at the shallowest / deepest sample of the block for this lane (t is monotone in the sample depth)
                    const float tlo = fmaf(smin, r, nm), thi = fmaf(smax, r, nm);
                    if (WIN)
                    {
                        if (__all_sync(0xffffffffu, tlo >= tsat)) { base = fmaf(A, esat, base); sat += n_real; continue; }
                        if (__all_sync(0xffffffffu, thi <= -tsat)) { base = fmaf(-A, esat, base); sat += n_real; continue; }
                    }
                    exec += n_real;
                    // Sign-uniform occluder: every sample of the block lies behind it (all t >= 0) or in front of it (all t <= 0)
                    // for every lane -- the common case once the list is depth-sorted.  erf(t) = +-(1 - w(t)) with the sign known
                    // per (occluder, block): the +-A goes to `base` once, each term only accumulates -+A w(t), which drops the
                    // per-term sign transfer (LOP3) and the 1 - w from the loop body (8 packed FMA-pipe ops + 1 MUFU per term).
                    const bool pos = __all_sync(0xffffffffu, tlo >= 0.f);
                    const bool neg = !pos && __all_sync(0xffffffffu, thi <= 0.f);
                    if (PACK && (pos || neg))
                    {
                        base += pos ? A : -A;
                        const float sA = pos ? -A : A;
                        const float2 rr = make_float2(r, r), mm = make_float2(nm, nm), AA = make_float2(sA, sA);
#pragma unroll
                        for (int e2 = 0; e2 < Q / 2; ++e2)
#pragma unroll
                            for (int k = 0; k < 5; ++k)
                            {
                                const float2 t = __ffma2_rn(make_float2(s[2 * e2][k], s[2 * e2 + 1][k]), rr, mm);
                                const float2 ac = __ffma2_rn(AA, erfc_mag2<ERF>(t), make_float2(acc[2 * e2][k], acc[2 * e2 + 1][k]));
                                acc[2 * e2][k] = ac.x;
                                acc[2 * e2 + 1][k] = ac.y;
                            }
                        continue;
                    }
                    if (PACK)
                    {
                        const float2 rr = make_float2(r, r), mm = make_float2(nm, nm), AA = make_float2(A, A);
#pragma unroll
                        for (int e2 = 0; e2 < Q / 2; ++e2)
#pragma unroll
                            for (int k = 0; k < 5; ++k)
                            {
                                const float2 t = __ffma2_rn(make_float2(s[2 * e2][k], s[2 * e2 + 1][k]), rr, mm);
                                const float2 ev = erf_variant2<ERF>(t);
                                const float2 ac = __ffma2_rn(AA, ev, make_float2(acc[2 * e2][k], acc[2 * e2 + 1][k]));
                                acc[2 * e2][k] = ac.x;
                                acc[2 * e2 + 1][k] = ac.y;
                            }
                        if (Q & 1)
                        {
#pragma unroll
                            for (int k = 0; k < 5; ++k) acc[Q - 1][k] = fmaf(A, erf_variant<ERF>(fmaf(s[Q - 1][k], r, nm)), acc[Q - 1][k]);
                        }
                    }
                    else
                    {
#pragma unroll
                        for (int e = 0; e < Q; ++e)
#pragma unroll
                            for (int k = 0; k < 5; ++k) acc[e][k] = fmaf(A, erf_variant<ERF>(fmaf(s[e][k], r, nm)), acc[e][k]);
                    }
                }
                if (!resident) chunk_end(c);
            }
            // T(s) = 2^(C - acc); pdf at the samples = c_bar e^{-k^2/2}, k = -4..0 (src/vrt/rt.h:153-161)
#pragma unroll
            for (int e = 0; e < Q; ++e)
            {
                const float Cb = C - base;
                float inner = 3.3546262790251185e-4f * ex2_approx(Cb - acc[e][0]);
                inner = fmaf(1.1108996538242306e-2f, ex2_approx(Cb - acc[e][1]), inner);
                inner = fmaf(1.3533528323661270e-1f, ex2_approx(Cb - acc[e][2]), inner);
                inner = fmaf(6.0653065971263342e-1f, ex2_approx(Cb - acc[e][3]), inner);
                inner += ex2_approx(Cb - acc[e][4]);
                inner *= wgt[e];
                Lr = fmaf(alb[e].x, inner, Lr);
                Lg = fmaf(alb[e].y, inner, Lg);
                Lb = fmaf(alb[e].z, inner, Lb);
                La = fmaf(alb[e].w, inner, La);
            }
        }

        // ---- K3: framebuffer ----
        if (slot != NO_SLOT) args.partial[(size_t)(slot + slice) * 32 + lane] = make_float4(Lr, Lg, Lb, La); // summed by k3_combine
        else if (live) store_pixel(args, (size_t)py * G.W + px, Lr, Lg, Lb, La);
        if (lane == 0 && exec) atomicAdd(args.terms_exec, exec * 5ull * n_live);
        if (WIN && lane == 0 && sat) atomicAdd(args.terms_sat, sat * 5ull * n_live);
        exec = 0;
        sat = 0;
    }
}

// K3', split cells: sum the slices' partial radiances in slice order (deterministic) and write the pixel.
__global__ void __launch_bounds__(256) k3_combine(const RenderArgs args, const uint32_t *__restrict__ list_off, int cy_begin, int cy_end)
{
    const FrameGeom &G = c_geom;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t ncells = (uint32_t)((cy_end - cy_begin) * G.ncx);
    if (w >= ncells) return;
    const int cx = (int)(w % G.ncx), cy = cy_begin + (int)(w / G.ncx);
    const uint32_t cell = (uint32_t)(cy * G.ncx + cx);
    int x0, y0, cw, ch;
    cell_rect(cx, cy, x0, y0, cw, ch);
    if (min(y0 + ch, G.row_end) <= max(y0, G.row_begin)) return; // not queued: its slot entry is stale
    const uint32_t slot = args.cell_slot[cell];
    if (slot == NO_SLOT) return;
    const uint32_t lid = cell_list_id(cx, cy);
    const uint32_t items = cell_items(list_off[lid + 1] - list_off[lid], cell);
    float4 L = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t k = 0; k < items; ++k)
    {
        const float4 p = args.partial[(size_t)(slot + k) * 32 + lane];
        L.x += p.x; L.y += p.y; L.z += p.z; L.w += p.w;
    }
    const int lx = lane & (CELL_W - 1), ly = lane >> 3;
    const int px = x0 + lx, py = y0 + ly;
    if (lx < cw && ly < ch && py >= G.row_begin && py < G.row_end) store_pixel(args, (size_t)py * G.W + px, L.x, L.y, L.z, L.w);
}

// ------------------------------------------------------------------------------------------------
// K2'', the reference's alternative approximations as selectable device functions (VRT_CUDA_APPROX_*)
// ------------------------------------------------------------------------------------------------
// spline_erf / spline_erf_mirror / taylor_erf and fast_exp / spline_exp of src/vrt/approx.cpp plugged into the same
// hoisted sums as k2_render, so the variant comparison of tests/img-error.cpp and the tables of tests/accuracy.cpp run on the
// GPU.  The approximations are not odd, not monotone and (the splines) not even continuous, so none of k2_render's
// saturation / sign shortcuts apply: this kernel evaluates every term with the selected functions, in natural-log units.
struct ApproxTables
{
    float4 erf_coef[VRT_SPLINE_ERF_SEGMENTS], exp_coef[VRT_SPLINE_EXP_SEGMENTS];
    float erf_knot[VRT_SPLINE_ERF_SEGMENTS + 1], exp_knot[VRT_SPLINE_EXP_SEGMENTS + 1];
};
__constant__ ApproxTables c_approx;

enum { ERFV_AS = 0, ERFV_EXACT = 1, ERFV_SPLINE = 2, ERFV_SPLINE_MIRROR = 3, ERFV_TAYLOR = 4 };
enum { EXPV_EXACT = 0, EXPV_FAST = 1, EXPV_SPLINE = 2 };

// cubic of the segment [knot[i], knot[i+1]) that holds x among the first NSEG segments (x below knot[1] -> segment 0, at or
// above knot[NSEG-1] -> the last one); the knots are read warp-uniformly, the coefficients per lane
template <int NSEG>
__device__ __forceinline__ float spline_segment(const float *knot, const float4 *coef, float x)
{
    int i = 0;
#pragma unroll
    for (int k = 1; k < NSEG; ++k) i += (x >= knot[k]) ? 1 : 0;
    const float4 c = coef[i];
    const float d = x - knot[i];
    return fmaf(fmaf(fmaf(c.x, d, c.y), d, c.z), d, c.w);
}

template <int ERFV>
__device__ __forceinline__ float erf_approx(const ApproxTables &T, float t)
{
    if (ERFV == ERFV_AS) return erf_as(t);
    if (ERFV == ERFV_EXACT) return erf_exact(t);
    if (ERFV == ERFV_SPLINE)
    {
        // src/vrt/approx.cpp:9-23: -1 up to the first knot, +1 from the last one on
        const float v = spline_segment<VRT_SPLINE_ERF_SEGMENTS>(T.erf_knot, T.erf_coef, t);
        return t <= T.erf_knot[0] ? -1.f : (t >= T.erf_knot[VRT_SPLINE_ERF_SEGMENTS] ? 1.f : v);
    }
    if (ERFV == ERFV_SPLINE_MIRROR)
    {
        // src/vrt/approx.cpp:45-56: the negative half (segments 0..3, then segment 4 all the way to 0) at -|t|, mirrored;
        // sign(0) = +1
        const float m = -fabsf(t);
        float v = spline_segment<5>(T.erf_knot, T.erf_coef, m);
        v = m <= T.erf_knot[0] ? -1.f : v;
        return t >= 0.f ? -v : v;
    }
    // src/vrt/approx.cpp:64-77: ten Maclaurin terms (-1)^n / (n! (2n+1)), saturated outside (-2, 2)
    const float x2 = t * t;
    float p = -1.f / 6894720.f;
    p = fmaf(p, x2, 1.f / 685440.f);
    p = fmaf(p, x2, -1.f / 75600.f);
    p = fmaf(p, x2, 1.f / 9360.f);
    p = fmaf(p, x2, -1.f / 1320.f);
    p = fmaf(p, x2, 1.f / 216.f);
    p = fmaf(p, x2, -1.f / 42.f);
    p = fmaf(p, x2, 1.f / 10.f);
    p = fmaf(p, x2, -1.f / 3.f);
    p = fmaf(p, x2, 1.f);
    const float v = (2.f * 0.5641895835477563f) * p * t;
    return t <= -2.f ? -1.f : (t >= 2.f ? 1.f : v);
}

template <int EXPV>
__device__ __forceinline__ float exp_approx(const ApproxTables &T, float x)
{
    if (EXPV == EXPV_EXACT) return ex2_approx(x * LOG2E);
    if (EXPV == EXPV_FAST)
    {
        // src/vrt/approx.cpp:112-137 (Schraudolph): the integer nearest to a x + b is the bit pattern of the result.  Range
        // clamp as in the reference's non-NDEBUG build (the conversion is undefined outside it); rounding as simd::cvts.
        constexpr float a = 8388608.f / 0.6931471805599453f, b = 8388608.f * (127.f - 0.043677448f);
        float y = fmaf(a, x, b);
        y = y < 8388608.f ? 0.f : fminf(y, 8388608.f * 255.f);
        return __uint_as_float((uint32_t)__float2int_rn(y));
    }
    // src/vrt/approx.cpp:141-163: 0 up to the first knot, 1 from the last one (x = 0) on
    const float v = spline_segment<VRT_SPLINE_EXP_SEGMENTS>(T.exp_knot, T.exp_coef, x);
    return x <= T.exp_knot[0] ? 0.f : (x >= T.exp_knot[VRT_SPLINE_EXP_SEGMENTS] ? 1.f : v);
}

__device__ __forceinline__ void load_tables(ApproxTables &s_tab)
{
    const float *src = reinterpret_cast<const float *>(&c_approx);
    float *dst = reinterpret_cast<float *>(&s_tab);
    for (uint32_t i = threadIdx.x; i < sizeof(ApproxTables) / sizeof(float); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

constexpr int VQ = 4; // emitters per register block of the variant kernel

template <int ERFV, int EXPV>
__global__ void __launch_bounds__(K2_WARPS * 32) k2_variant(const RenderArgs args)
{
    __shared__ ApproxTables s_tab;
    load_tables(s_tab);
    const FrameGeom &G = c_geom;
    const int lane = threadIdx.x & 31;
    const int lx = lane & (CELL_W - 1), ly = lane >> 3;
    constexpr float LN2 = 1.f / LOG2E;
    for (;;)
    {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(args.counter, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= args.n_queue) break;
        const uint32_t item = args.queue[qi];
        const uint32_t cell = item & ((1u << ITEM_CELL_BITS) - 1u), slice = item >> ITEM_CELL_BITS;
        const int cx = cell % G.ncx, cy = cell / G.ncx;
        int x0, y0, cw, ch;
        cell_rect(cx, cy, x0, y0, cw, ch);
        const int px = x0 + min(lx, cw - 1), py = y0 + min(ly, ch - 1);
        const bool live = lx < cw && ly < ch && py >= G.row_begin && py < G.row_end;
        const uint32_t n_live = __popc(__ballot_sync(0xffffffffu, live));
        const PixelRay ray = pixel_ray(px, py);
        const uint32_t lid = cell_list_id(cx, cy);
        const uint32_t off = args.list_off[lid];
        const uint32_t n = args.list_off[lid + 1] - off;
        auto load_rec = [&](uint32_t k) -> const Rec * { return args.rec + (args.list_idx ? args.list_idx[off + k] : off + k); };
        // occluder j for this lane: mu_bar, weight A = sigma c sqrt(pi/2) Exp(-d^2 / 2 sigma^2), r = 1/(sqrt2 sigma)
        auto occluder = [&](const Rec *rc, float &mu, float &A, float &r) {
            const float4 a = rc->a, b = rc->b;
            mu = fmaf(a.z, ray.nz, fmaf(a.y, ray.ny, a.x * ray.nx));
            const float qx = fmaf(-mu, ray.nx, a.x), qy = fmaf(-mu, ray.ny, a.y), qz = fmaf(-mu, ray.nz, a.z);
            const float d2 = fmaf(qz, qz, fmaf(qy, qy, fmaf(qx, qx, a.w)));
            A = (b.z * LN2) * exp_approx<EXPV>(s_tab, -d2 * (b.y * LN2));
            r = b.x;
        };

        // pass A: C = sum_j A_j Erf(-m_j)
        float C = 0.f;
        for (uint32_t j = 0; j < n; ++j)
        {
            float mu, A, r;
            occluder(load_rec(j), mu, A, r);
            C = fmaf(A, erf_approx<ERFV>(s_tab, -mu * r), C);
        }

        // pass B
        float Lr = 0.f, Lg = 0.f, Lb = 0.f, La = 0.f;
        unsigned long long exec = 0;
        const uint32_t slot = args.cell_slot ? args.cell_slot[cell] : NO_SLOT;
        const uint32_t q_begin = slot != NO_SLOT ? slice * (uint32_t)G.slice : 0u, q_end = slot != NO_SLOT ? min(n, q_begin + (uint32_t)G.slice) : n;
        for (uint32_t q0 = q_begin; q0 < q_end; q0 += VQ)
        {
            float s[VQ][5], acc[VQ][5], wgt[VQ];
            float4 alb[VQ];
            float s0 = 0.f;
#pragma unroll
            for (int e = 0; e < VQ; ++e)
            {
                const bool real = q0 + e < q_end;
                const Rec *rc = load_rec(real ? q0 + e : q0);
                const float4 b = rc->b;
                alb[e] = rc->c;
                float mu, ee;
                occluder_setup(rc->a, b, ray, mu, ee);
                if (e == 0)
                {
                    s0 = __shfl_sync(0xffffffffu, mu, 0);
                    s0 = (fabsf(s0) <= 3.0e38f) ? s0 : 0.f;
                }
                // the density G_q at the samples always uses the exact exp (types.h:204-208; template default of the SIMD pdf)
                wgt[e] = real ? b.z * ee * (1.f / (SQRT_PI_2 * LOG2E)) : 0.f;
#pragma unroll
                for (int k = 0; k < 5; ++k)
                {
                    s[e][k] = (mu - s0) + (float)(k - 4) * b.w;
                    acc[e][k] = 0.f;
                }
            }
            const uint32_t n_real = min((uint32_t)VQ, q_end - q0);
            for (uint32_t j = 0; j < n; ++j)
            {
                float mu, A, r;
                occluder(load_rec(j), mu, A, r);
                // a weight of exactly 0 for the whole warp contributes exactly 0 to every sum
                if (args.skip_thresh >= 0.f && !__any_sync(0xffffffffu, A != 0.f)) continue;
                exec += n_real;
                const float nm = -(mu - s0) * r;
#pragma unroll
                for (int e = 0; e < VQ; ++e)
#pragma unroll
                    for (int k = 0; k < 5; ++k)
                    {
                        // two roundings, not an FMA: an emitter's own k = 0 sample must give t = 0 EXACTLY, as the reference's
                        // s/(sqrt2 sigma) - mu_bar/(sqrt2 sigma) does -- spline_erf_mirror jumps by 0.107 across t = 0
                        const float t = __fadd_rn(__fmul_rn(s[e][k], r), nm);
                        acc[e][k] = fmaf(A, erf_approx<ERFV>(s_tab, t), acc[e][k]);
                    }
            }
#pragma unroll
            for (int e = 0; e < VQ; ++e)
            {
                float inner = 3.3546262790251185e-4f * exp_approx<EXPV>(s_tab, C - acc[e][0]);
                inner = fmaf(1.1108996538242306e-2f, exp_approx<EXPV>(s_tab, C - acc[e][1]), inner);
                inner = fmaf(1.3533528323661270e-1f, exp_approx<EXPV>(s_tab, C - acc[e][2]), inner);
                inner = fmaf(6.0653065971263342e-1f, exp_approx<EXPV>(s_tab, C - acc[e][3]), inner);
                inner += exp_approx<EXPV>(s_tab, C - acc[e][4]);
                inner *= wgt[e];
                Lr = fmaf(alb[e].x, inner, Lr);
                Lg = fmaf(alb[e].y, inner, Lg);
                Lb = fmaf(alb[e].z, inner, Lb);
                La = fmaf(alb[e].w, inner, La);
            }
        }
        if (slot != NO_SLOT) args.partial[(size_t)(slot + slice) * 32 + lane] = make_float4(Lr, Lg, Lb, La);
        else if (live) store_pixel(args, (size_t)py * G.W + px, Lr, Lg, Lb, La);
        if (lane == 0 && exec) atomicAdd(args.terms_exec, exec * 5ull * n_live);
    }
}

// the functions tests/accuracy.cpp tabulates, evaluated on the device
__global__ void k_approx_table(int fn, const float *__restrict__ x, float *__restrict__ y, uint64_t n)
{
    __shared__ ApproxTables s_tab;
    load_tables(s_tab);
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    float r;
    switch (fn)
    {
    case VRT_CUDA_FN_SPLINE_ERF: r = erf_approx<ERFV_SPLINE>(s_tab, v); break;
    case VRT_CUDA_FN_SPLINE_ERF_MIRROR: r = erf_approx<ERFV_SPLINE_MIRROR>(s_tab, v); break;
    case VRT_CUDA_FN_TAYLOR_ERF: r = erf_approx<ERFV_TAYLOR>(s_tab, v); break;
    case VRT_CUDA_FN_AS_ERF: r = erf_approx<ERFV_AS>(s_tab, v); break;
    case VRT_CUDA_FN_ERF: r = erf_approx<ERFV_EXACT>(s_tab, v); break;
    case VRT_CUDA_FN_EXP: r = exp_approx<EXPV_EXACT>(s_tab, v); break;
    case VRT_CUDA_FN_FAST_EXP: r = exp_approx<EXPV_FAST>(s_tab, v); break;
    default: r = exp_approx<EXPV_SPLINE>(s_tab, v); break;
    }
    y[i] = r;
}

// ------------------------------------------------------------------------------------------------
// K2', depth-window render (VRT_CUDA_DEPTH_WINDOW) for cells whose list fits the per-warp cache
// ------------------------------------------------------------------------------------------------
// Lists are depth-sorted by K1.  Pass A walks the list once per pixel, accumulates C and stores the per-lane PREFIX SUMS of
// the weights A_j in shared memory, plus two warp-uniform depths per occluder: beyond f_j every lane's erf argument is
// >= t_sat (the occluder is entirely in front: erf = +esat), before b_j it is <= -t_sat (entirely behind: -esat).  For an
// emitter block whose samples span [Smin, Smax] the leading occluders with f_j <= Smin and the trailing ones with
// b_j >= Smax are resolved together as  esat (P[f] - (P[n] - P[b]))  -- two shared-memory reads -- and only the window
// [f, b) in between is evaluated term by term.
constexpr int WIN_Q = 8;
struct WinSmem
{
    float prefix[WIN_CAP + 1][32]; // prefix[j][lane] = sum_{i<j} A_i(lane), log2 units
    float4 a[WIN_CAP], b[WIN_CAP]; // occluder part of the records
    float2 fb[WIN_CAP];            // (f_j, b_j)
};

__device__ __forceinline__ int ordered_int(float x)
{
    const int k = __float_as_int(x);
    return k ^ ((k >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ordered_float(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
__device__ __forceinline__ float warp_max_f(float x) { return ordered_float(__reduce_max_sync(0xffffffffu, ordered_int(x))); }
__device__ __forceinline__ float warp_min_f(float x) { return ordered_float(__reduce_min_sync(0xffffffffu, ordered_int(x))); }

template <int ERF>
__global__ void __launch_bounds__(K2_WARPS * 32, 1) k2_window(const RenderArgs args, uint32_t queue_begin)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    constexpr int Q = WIN_Q;
    const FrameGeom &G = c_geom;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WinSmem &sm = reinterpret_cast<WinSmem *>(s_raw)[warp];
    const int lx = lane & (CELL_W - 1), ly = lane >> 3;
    const float tsat = ERF == 0 ? 5.5f : EX_XMAX;
    const float esat = erf_variant<ERF>(tsat);

    for (;;)
    {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(args.counter + 1, 1u) + queue_begin;
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= args.n_queue) break;
        const uint32_t item = args.queue[qi];
        const uint32_t cell = item & ((1u << ITEM_CELL_BITS) - 1u), slice = item >> ITEM_CELL_BITS;
        const int cx = cell % G.ncx, cy = cell / G.ncx;
        int x0, y0, cw, ch;
        cell_rect(cx, cy, x0, y0, cw, ch);
        const int px = x0 + min(lx, cw - 1), py = y0 + min(ly, ch - 1);
        const bool live = lx < cw && ly < ch && py >= G.row_begin && py < G.row_end;
        const uint32_t n_live = __popc(__ballot_sync(0xffffffffu, live));
        const PixelRay ray = pixel_ray(px, py);
        const uint32_t lid = cell_list_id(cx, cy);
        const uint32_t off = args.list_off[lid];
        const uint32_t n = min(args.list_off[lid + 1] - off, (uint32_t)WIN_CAP); // (longer lists never reach this kernel)

        // stage the whole list (occluder part) once
        __syncwarp();
        for (uint32_t j = lane; j < n; j += 32)
        {
            const Rec *r = args.rec + args.list_idx[off + j];
            sm.a[j] = r->a;
            sm.b[j] = r->b;
        }
        __syncwarp();

        // ---- pass A: C, prefix sums of the weights, saturation depths ----
        float C = 0.f, run = 0.f;
        sm.prefix[0][lane] = 0.f;
        for (uint32_t j = 0; j < n; ++j)
        {
            const float4 a = sm.a[j], b = sm.b[j];
            float mu, e;
            occluder_setup(a, b, ray, mu, e);
            // an occluder no lane sees (weight <= threshold everywhere) is dropped exactly like the plain kernel's skip
            const bool alive = __any_sync(0xffffffffu, e > args.skip_thresh);
            const float A = alive ? b.z * e : 0.f;
            C = fmaf(A, erf_variant<ERF>(-mu * b.x), C);
            run += A;
            sm.prefix[j + 1][lane] = run;
            const float mumax = warp_max_f(mu), mumin = warp_min_f(mu);
            if (lane == 0)
            {
                const float half = tsat * 1.0000005f / b.x + 1e-6f * fabsf(mumax); // t >= tsat must hold after fp32 rounding of t
                sm.fb[j] = alive ? make_float2(mumax + half, mumin - half) : make_float2(-3.0e38f, 3.0e38f);
            }
        }
        __syncwarp();
        const float total = run;

        // ---- pass B ----
        float Lr = 0.f, Lg = 0.f, Lb = 0.f, La = 0.f;
        unsigned long long exec = 0, sat = 0;
        // a split cell's item covers the emitters [q_begin, q_end) only (pass A above is per item)
        const uint32_t slot = args.cell_slot ? args.cell_slot[cell] : NO_SLOT;
        const uint32_t q_begin = slot != NO_SLOT ? slice * (uint32_t)G.slice : 0u, q_end = slot != NO_SLOT ? min(n, q_begin + (uint32_t)G.slice) : n;
        for (uint32_t q0 = q_begin; q0 < q_end; q0 += Q)
        {
            float s[Q][5], acc[Q][5], wgt[Q];
            float4 alb[Q];
            float s0 = 0.f, smin = 3.0e38f, smax = -3.0e38f;
            bool any_emit = false;
#pragma unroll
            for (int e = 0; e < Q; ++e)
            {
                const bool real = q0 + e < q_end;
                const uint32_t je = real ? q0 + e : q0;
                const float4 a = sm.a[je], b = sm.b[je];
                alb[e] = args.rec[args.list_idx[off + je]].c;
                float mu, ee;
                occluder_setup(a, b, ray, mu, ee);
                if (e == 0)
                {
                    s0 = __shfl_sync(0xffffffffu, mu, 0);
                    s0 = (fabsf(s0) <= 3.0e38f) ? s0 : 0.f;
                }
                wgt[e] = real ? b.z * ee * (1.f / (SQRT_PI_2 * LOG2E)) : 0.f;
                any_emit |= real && (ee > args.skip_thresh);
                if (real)
                {
                    smin = fminf(smin, mu - 4.f * b.w);
                    smax = fmaxf(smax, mu);
                }
#pragma unroll
                for (int k = 0; k < 5; ++k)
                {
                    s[e][k] = (mu - s0) + (float)(k - 4) * b.w;
                    acc[e][k] = 0.f;
                }
            }
            if (!__any_sync(0xffffffffu, any_emit)) continue;
            const uint32_t n_real = min((uint32_t)Q, q_end - q0);
            const float Smin = warp_min_f(smin), Smax = warp_max_f(smax);

            // leading run of occluders entirely in front of every sample, trailing run entirely behind
            uint32_t f = 0, bk = n;
            for (uint32_t j0 = 0; j0 < n; j0 += 32)
            {
                const uint32_t j = j0 + lane;
                const uint32_t m = __ballot_sync(0xffffffffu, j < n && sm.fb[j].x <= Smin);
                if (m == 0xffffffffu) { f = j0 + 32; continue; }
                f = j0 + (uint32_t)__ffs(~m) - 1u;
                break;
            }
            f = min(f, n);
            for (int j0 = (int)((n - 1) & ~31u); j0 >= 0; j0 -= 32)
            {
                const uint32_t j = (uint32_t)j0 + lane;
                // lanes beyond the list count as "behind" so the trailing run can start at the list end
                const uint32_t m = __ballot_sync(0xffffffffu, j >= n || sm.fb[j].y >= Smax);
                if (m == 0xffffffffu) { bk = (uint32_t)j0; continue; }
                bk = (uint32_t)j0 + 32u - (uint32_t)__clz(~m);
                break;
            }
            bk = max(bk, f);
            float base = esat * (sm.prefix[f][lane] - (total - sm.prefix[bk][lane]));
            sat += (unsigned long long)(f + (n - bk)) * n_real;
            const float smin0 = smin - s0, smax0 = smax - s0;

            for (uint32_t j = f; j < bk; ++j)
            {
                const float4 a = sm.a[j], b = sm.b[j];
                float mu, e;
                occluder_setup(a, b, ray, mu, e);
                if (!__any_sync(0xffffffffu, e > args.skip_thresh)) continue;
                const float A = b.z * e, r = b.x, nm = -(mu - s0) * r;
                exec += n_real;
                // sign-uniform occluder (see k2_render): +-A once, -+A w(t) per term
                const bool pos = __all_sync(0xffffffffu, fmaf(smin0, r, nm) >= 0.f);
                const bool neg = !pos && __all_sync(0xffffffffu, fmaf(smax0, r, nm) <= 0.f);
                if (pos || neg)
                {
                    base += pos ? A : -A;
                    const float sA = pos ? -A : A;
                    const float2 rr = make_float2(r, r), mm = make_float2(nm, nm), AA = make_float2(sA, sA);
#pragma unroll
                    for (int e2 = 0; e2 < Q / 2; ++e2)
#pragma unroll
                        for (int k = 0; k < 5; ++k)
                        {
                            const float2 t = __ffma2_rn(make_float2(s[2 * e2][k], s[2 * e2 + 1][k]), rr, mm);
                            const float2 ac = __ffma2_rn(AA, erfc_mag2<ERF>(t), make_float2(acc[2 * e2][k], acc[2 * e2 + 1][k]));
                            acc[2 * e2][k] = ac.x;
                            acc[2 * e2 + 1][k] = ac.y;
                        }
                    continue;
                }
                const float2 rr = make_float2(r, r), mm = make_float2(nm, nm), AA = make_float2(A, A);
#pragma unroll
                for (int e2 = 0; e2 < Q / 2; ++e2)
#pragma unroll
                    for (int k = 0; k < 5; ++k)
                    {
                        const float2 t = __ffma2_rn(make_float2(s[2 * e2][k], s[2 * e2 + 1][k]), rr, mm);
                        const float2 ev = erf_variant2<ERF>(t);
                        const float2 ac = __ffma2_rn(AA, ev, make_float2(acc[2 * e2][k], acc[2 * e2 + 1][k]));
                        acc[2 * e2][k] = ac.x;
                        acc[2 * e2 + 1][k] = ac.y;
                    }
            }
            const float Cb = C - base;
#pragma unroll
            for (int e = 0; e < Q; ++e)
            {
                float inner = 3.3546262790251185e-4f * ex2_approx(Cb - acc[e][0]);
                inner = fmaf(1.1108996538242306e-2f, ex2_approx(Cb - acc[e][1]), inner);
                inner = fmaf(1.3533528323661270e-1f, ex2_approx(Cb - acc[e][2]), inner);
                inner = fmaf(6.0653065971263342e-1f, ex2_approx(Cb - acc[e][3]), inner);
                inner += ex2_approx(Cb - acc[e][4]);
                inner *= wgt[e];
                Lr = fmaf(alb[e].x, inner, Lr);
                Lg = fmaf(alb[e].y, inner, Lg);
                Lb = fmaf(alb[e].z, inner, Lb);
                La = fmaf(alb[e].w, inner, La);
            }
        }
        if (slot != NO_SLOT) args.partial[(size_t)(slot + slice) * 32 + lane] = make_float4(Lr, Lg, Lb, La); // summed by k3_combine
        else if (live) store_pixel(args, (size_t)py * G.W + px, Lr, Lg, Lb, La);
        if (lane == 0 && exec) atomicAdd(args.terms_exec, exec * 5ull * n_live);
        if (lane == 0 && sat) atomicAdd(args.terms_sat, sat * 5ull * n_live);
    }
}

// ------------------------------------------------------------------------------------------------
// FP32 roofline probe: dependent-free FFMA (or FFMA2) chains, the denominator of roofline.frac measured live
// ------------------------------------------------------------------------------------------------
template <bool PACK>
__global__ void __launch_bounds__(256) k_fp32_peak(float *out, int iters, float a, float b)
{
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-4f - i);
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            if (PACK) v[i] = __ffma2_rn(v[i], aa, bb);
            else
            {
                v[i].x = fmaf(v[i].x, a, b);
                v[i].y = fmaf(v[i].y, a, b);
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].y;
    if (acc == 12345.678f) out[0] = acc; // never true; keeps the chains alive
}

// Inner-term ceiling probe: the exact instruction mix of K2's body (per pair of terms 7 FFMA2 + 2 FMUL2 + 2 MUFU.RCP +
// 2 LOP3) with NP independent pairs per thread and no loads, setup or control flow around it.
template <int NP, bool SIGN_FREE>
__global__ void __launch_bounds__(256) k_term_peak(float *out, int iters, float r0, float nm0, float a0)
{
    float2 s[NP], acc[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i)
    {
        s[i] = make_float2(threadIdx.x * 1e-3f + i * 0.37f, blockIdx.x * 1e-4f - i * 0.21f);
        acc[i] = make_float2(0.f, 0.f);
    }
    float r = r0, nm = nm0, A = a0;
    for (int it = 0; it < iters; ++it)
    {
        const float2 rr = make_float2(r, r), mm = make_float2(nm, nm), AA = make_float2(A, A);
#pragma unroll
        for (int i = 0; i < NP; ++i)
            acc[i] = SIGN_FREE ? __ffma2_rn(AA, erfc_mag2<0>(__ffma2_rn(s[i], rr, mm)), acc[i]) : __ffma2_rn(AA, erf_variant2<0>(__ffma2_rn(s[i], rr, mm)), acc[i]);
        r += 1e-4f; nm -= 1e-4f; A += 1e-6f;
    }
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) t += acc[i].x + acc[i].y;
    if (t == 12345.678f) out[0] = t;
}

// Pipe-mix probe: NF packed FMAs + NM MUFU.RCP + NL LOP3 per step on 16 independent float2 chains per thread.
template <int NF, int NM, int NL>
__global__ void __launch_bounds__(256) k_mix_peak(float *out, int iters, float a, float b)
{
    float2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = make_float2(1.f + threadIdx.x * 1e-3f + i, 2.f + blockIdx.x * 1e-4f + i);
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 16; ++i)
        {
#pragma unroll
            for (int f = 0; f < NF; ++f) v[i] = __ffma2_rn(v[i], aa, bb);
            if (NM >= 1) v[i].x = rcp_approx(v[i].x);
            if (NM >= 2) v[i].y = rcp_approx(v[i].y);
            if (NL >= 1) v[i].x = copysign_bits(v[i].x, v[i].y);
            if (NL >= 2) v[i].y = copysign_bits(v[i].y, aa.x);
        }
    }
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += v[i].x + v[i].y;
    if (t == 12345.678f) out[0] = t;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct DevBuf
{
    void *p = nullptr;
    size_t cap = 0;
};

std::string g_create_error;
} // namespace

struct vrt_cuda_ctx
{
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;
    int sm_count = 148;

    uint64_t n_gauss = 0;
    DevBuf aos;        // scene, n x 10 floats
    DevBuf rec;        // frame records
    DevBuf cullrec;    // reference tiling projection
    DevBuf lvl_counts[4], lvl_offsets[4], lvl_idx[4], lvl_group_off; // coarse culling levels
    DevBuf ccounts, coffsets, cidx;       // level 1 (cells / tiles)
    DevBuf hist, queue, stats, counter, rowcost, scan_tmp, cell_slot, partial;
    DevBuf out_image, out_rad;
    DevBuf tile_aos; // host-supplied tile lists (concatenated)
    DevBuf tile_off;
    DevBuf lit_offsets, lit_idx; // literal lists (tile_gaussians membership) kept for vrt_cuda_get_lists while K2 uses the visible lists

    // state of the last tile()
    bool have_lists = false;
    bool lists_from_host = false;
    bool lists_sorted = false;
    uint32_t n_big = 0, n_split = 0;
    bool win_attr[2] = {false, false};
    FrameGeom geom{};
    uint32_t n_lists = 0;
    uint64_t n_entries = 0;
    uint32_t n_queue = 0;
    int cy_begin = 0, cy_end = 0;
    uint32_t launches = 0, render_launches = 0; // kernels of the last tile() / render()
    // literal list modes with culling: the lists K2 walks are "literal membership AND visible from the cell" (see
    // visible_pass); what the literal lists were is kept here
    bool literal = false;
    int lit_kind = 0;
    uint32_t lit_n_lists = 0;
    uint64_t lit_n_entries = 0;
    TileStats lit_stats{};
    float ms_tile = 0.f;
    // tuning
    int tune_q = 0; // 0 = automatic: 8 emitters per register block, 4 when the lists are short
    int tune_pack = 1;
};

namespace
{
int fail(vrt_cuda_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    return code;
}

#define CU(call)                                                                                                        \
    do                                                                                                                  \
    {                                                                                                                   \
        cudaError_t e_ = (call);                                                                                        \
        if (e_ != cudaSuccess) return fail(ctx, VRT_CUDA_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int reserve(vrt_cuda_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return 0;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    const size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) return fail(ctx, VRT_CUDA_E_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    b.cap = want;
    return 0;
}

void inverse4(const float *m, float *r)
{
    // adjugate / determinant, column-major
    const float a00 = m[0], a01 = m[1], a02 = m[2], a03 = m[3], a10 = m[4], a11 = m[5], a12 = m[6], a13 = m[7];
    const float a20 = m[8], a21 = m[9], a22 = m[10], a23 = m[11], a30 = m[12], a31 = m[13], a32 = m[14], a33 = m[15];
    const float b00 = a00 * a11 - a01 * a10, b01 = a00 * a12 - a02 * a10, b02 = a00 * a13 - a03 * a10;
    const float b03 = a01 * a12 - a02 * a11, b04 = a01 * a13 - a03 * a11, b05 = a02 * a13 - a03 * a12;
    const float b06 = a20 * a31 - a21 * a30, b07 = a20 * a32 - a22 * a30, b08 = a20 * a33 - a23 * a30;
    const float b09 = a21 * a32 - a22 * a31, b10 = a21 * a33 - a23 * a31, b11 = a22 * a33 - a23 * a32;
    const float id = 1.f / (b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06);
    r[0] = (a11 * b11 - a12 * b10 + a13 * b09) * id;  r[1] = (a02 * b10 - a01 * b11 - a03 * b09) * id;
    r[2] = (a31 * b05 - a32 * b04 + a33 * b03) * id;  r[3] = (a22 * b04 - a21 * b05 - a23 * b03) * id;
    r[4] = (a12 * b08 - a10 * b11 - a13 * b07) * id;  r[5] = (a00 * b11 - a02 * b08 + a03 * b07) * id;
    r[6] = (a32 * b02 - a30 * b05 - a33 * b01) * id;  r[7] = (a20 * b05 - a22 * b02 + a23 * b01) * id;
    r[8] = (a10 * b10 - a11 * b08 + a13 * b06) * id;  r[9] = (a01 * b08 - a00 * b10 - a03 * b06) * id;
    r[10] = (a30 * b04 - a31 * b02 + a33 * b00) * id; r[11] = (a21 * b02 - a20 * b04 - a23 * b00) * id;
    r[12] = (a11 * b07 - a10 * b09 - a12 * b06) * id; r[13] = (a00 * b09 - a01 * b07 + a02 * b06) * id;
    r[14] = (a31 * b01 - a30 * b03 - a32 * b00) * id; r[15] = (a20 * b03 - a21 * b01 + a22 * b00) * id;
}

// Validates the frame and fills the geometry; list_kind_override >= 0 forces the list kind (host tile lists).
int make_geom(vrt_cuda_ctx *ctx, const vrt_cuda_frame *f, int list_kind_override, FrameGeom &G, std::vector<float> &cxs, std::vector<float> &cys)
{
    if (!f) return fail(ctx, VRT_CUDA_E_INVALID, "frame is NULL");
    if (f->width == 0 || f->height == 0 || f->width > 65536 || f->height > 65536) return fail(ctx, VRT_CUDA_E_INVALID, "bad image size %ux%u", f->width, f->height);
    if (f->origin[3] != 0.f) return fail(ctx, VRT_CUDA_E_INVALID, "origin.w must be 0 (it is in every reference call site, main.cpp:248,253)");
    std::memset(&G, 0, sizeof(G));
    const uint32_t lm = f->flags & VRT_CUDA_LIST_MASK;
    const bool tiled = (list_kind_override == 1) || lm == VRT_CUDA_LIST_REFERENCE || lm == VRT_CUDA_LIST_REFERENCE_BOUND;
    G.W = (int)f->width;
    G.H = (int)f->height;
    G.tiles_x = tiled ? (int)f->tiles_x : 1;
    G.tiles_y = tiled ? (int)f->tiles_y : 1;
    if (tiled)
    {
        if (f->tiles_x == 0 || f->tiles_y == 0 || f->tiles_x > 1024 || f->tiles_y > 1024) return fail(ctx, VRT_CUDA_E_INVALID, "tiles per axis must be in 1..1024");
        if (f->width % f->tiles_x || f->height % f->tiles_y)
            return fail(ctx, VRT_CUDA_E_INVALID, "image %ux%u is not divisible into %ux%u tiles (the reference truncates tile_width, rt.h:254)", f->width, f->height, f->tiles_x, f->tiles_y);
    }
    G.tile_w = G.W / G.tiles_x;
    G.tile_h = G.H / G.tiles_y;
    G.cptx = (G.tile_w + CELL_W - 1) / CELL_W;
    G.cpty = (G.tile_h + CELL_H - 1) / CELL_H;
    G.ncx = G.tiles_x * G.cptx;
    G.ncy = G.tiles_y * G.cpty;
    G.uniform = (G.tile_w % CELL_W == 0 && G.tile_h % CELL_H == 0) ? 1 : 0;
    G.slice = SLICE_MAX;
    G.row_begin = (int)f->row_begin;
    G.row_end = (int)f->row_end;
    if (f->row_begin == 0 && f->row_end == 0) G.row_end = G.H;
    if (G.row_begin < 0 || G.row_end > G.H || G.row_begin >= G.row_end) return fail(ctx, VRT_CUDA_E_INVALID, "bad row band [%u,%u)", f->row_begin, f->row_end);
    G.use_ref = (lm == VRT_CUDA_LIST_REFERENCE || lm == VRT_CUDA_LIST_REFERENCE_BOUND) ? 1 : 0;
    G.use_bound = (lm == VRT_CUDA_LIST_REFERENCE_BOUND || lm == VRT_CUDA_LIST_BOUND) ? 1 : 0;
    G.list_kind = G.use_bound ? 0 : (G.use_ref ? 1 : 2);
    if (list_kind_override >= 0)
    {
        G.list_kind = list_kind_override;
        G.use_ref = G.use_bound = 0;
    }
    G.bound_k = f->bound_sigmas > 0.f ? f->bound_sigmas : 6.0f;
    G.tw = 2.f / (float)G.tiles_x;
    G.th = 2.f / (float)G.tiles_y;
    G.half_w = (float)G.W / 2.f;
    G.half_h = (float)G.H / 2.f;
    std::memcpy(G.view, f->view, sizeof(G.view));
    float inv[16];
    inverse4(f->view, inv);
    for (int i = 0; i < 3; ++i)
    {
        G.inv0[i] = inv[i];
        G.inv1[i] = inv[4 + i];
        G.inv3[i] = inv[12 + i];
        G.origin[i] = f->origin[i];
        if (!std::isfinite(inv[i]) || !std::isfinite(inv[4 + i]) || !std::isfinite(inv[12 + i])) return fail(ctx, VRT_CUDA_E_INVALID, "view matrix is singular");
    }
    // tile centres by the reference's float accumulation (src/vrt/rt.cpp:47-49)
    cxs.clear();
    cys.clear();
    for (float x = -1.f + G.tw / 2; x < 1.f; x += G.tw) cxs.push_back(x);
    for (float y = -1.f + G.th / 2; y < 1.f; y += G.th) cys.push_back(y);
    if (G.use_ref && ((int)cxs.size() != G.tiles_x || (int)cys.size() != G.tiles_y))
        return fail(ctx, VRT_CUDA_E_INVALID, "tile count %dx%d is not reproduced by the reference's float accumulation (%zu x %zu centres)", G.tiles_x, G.tiles_y, cxs.size(), cys.size());
    cxs.resize(1024, 0.f);
    cys.resize(1024, 0.f);
    return 0;
}

int upload_geom(vrt_cuda_ctx *ctx, const FrameGeom &G, const std::vector<float> &cxs, const std::vector<float> &cys)
{
    CU(cudaMemcpyToSymbolAsync(c_geom, &G, sizeof(G), 0, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyToSymbolAsync(c_tile_cx, cxs.data(), sizeof(float) * 1024, 0, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyToSymbolAsync(c_tile_cy, cys.data(), sizeof(float) * 1024, 0, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// exclusive scan of `n` counts into n + 1 offsets on the context's stream
int scan_u32(vrt_cuda_ctx *ctx, const uint32_t *counts, uint32_t *offsets, uint32_t n)
{
    if (n <= 4 * SCAN_TILE)
    {
        k1_scan<<<1, 1024, 0, ctx->stream>>>(counts, offsets, n);
        ctx->launches++;
        return 0;
    }
    const uint32_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (int rc = reserve(ctx, ctx->scan_tmp, sizeof(uint32_t) * (2 * (size_t)n_tiles + 2))) return rc;
    uint32_t *sums = (uint32_t *)ctx->scan_tmp.p, *offs = sums + n_tiles;
    k1_scan_tiles<<<n_tiles, 256, 0, ctx->stream>>>(counts, sums, n);
    k1_scan<<<1, 1024, 0, ctx->stream>>>(sums, offs, n_tiles);
    k1_scan_apply<<<n_tiles, 256, 0, ctx->stream>>>(counts, offs, offsets, n, n_tiles);
    ctx->launches += 3;
    return 0;
}

// cost ordering of the band's cells + listed-term statistics
int build_queue(vrt_cuda_ctx *ctx)
{
    const FrameGeom &G = ctx->geom;
    // cell rows intersecting the band
    int cyb = G.ncy, cye = 0;
    for (int cy = 0; cy < G.ncy; ++cy)
    {
        const int y0 = (cy / G.cpty) * G.tile_h + (cy % G.cpty) * CELL_H;
        const int h = std::min(CELL_H, G.tile_h - (cy % G.cpty) * CELL_H);
        if (y0 + h > G.row_begin && y0 < G.row_end)
        {
            cyb = std::min(cyb, cy);
            cye = std::max(cye, cy + 1);
        }
    }
    ctx->cy_begin = cyb;
    ctx->cy_end = cye;
    const int ncells = (cye - cyb) * G.ncx;
    if (int rc = reserve(ctx, ctx->hist, sizeof(uint32_t) * 65536)) return rc;
    if (int rc = reserve(ctx, ctx->stats, sizeof(TileStats))) return rc;
    if (int rc = reserve(ctx, ctx->counter, sizeof(uint32_t) * 4)) return rc;
    if (int rc = reserve(ctx, ctx->rowcost, sizeof(double) * (size_t)G.ncy)) return rc;
    CU(cudaMemsetAsync(ctx->hist.p, 0, sizeof(uint32_t) * 65536, ctx->stream));
    CU(cudaMemsetAsync(ctx->stats.p, 0, sizeof(TileStats), ctx->stream));
    CU(cudaMemsetAsync(ctx->rowcost.p, 0, sizeof(double) * (size_t)G.ncy, ctx->stream));
    const uint32_t *loff = (const uint32_t *)ctx->coffsets.p;
    const int tb = 256, gb = (ncells + tb - 1) / tb;
    k1_list_stats<<<(ctx->n_lists + 255) / 256, 256, 0, ctx->stream>>>(loff, ctx->n_lists, (TileStats *)ctx->stats.p);
    k1_hist<false><<<gb, tb, 0, ctx->stream>>>(loff, (uint32_t *)ctx->hist.p, (TileStats *)ctx->stats.p, (double *)ctx->rowcost.p, cyb, cye);
    TileStats ts;
    CU(cudaMemcpyAsync(&ts, ctx->stats.p, sizeof(ts), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    {
        // Slice size of split cells: as large as possible (every item repeats pass A), but no item may exceed a quarter of
        // the average work per resident warp, or the longest lists would decide the frame time on small frames.
        const double per_warp = ts.terms_listed / ((double)ctx->sm_count * 12.0);
        int slice = SLICE_MAX;
        while (slice > SLICE_MIN && 160.0 * slice * (double)ts.max_list > per_warp / 4.0) slice /= 2;
        ctx->geom.slice = slice;
        CU(cudaMemcpyToSymbolAsync(c_geom, &ctx->geom, sizeof(FrameGeom), 0, cudaMemcpyHostToDevice, ctx->stream));
    }
    k1_hist<true><<<gb, tb, 0, ctx->stream>>>(loff, (uint32_t *)ctx->hist.p, (TileStats *)ctx->stats.p, nullptr, cyb, cye);
    k1_hist_scan<<<1, 1024, 0, ctx->stream>>>((uint32_t *)ctx->hist.p);
    // descending order: the start slot of key WIN_CAP = number of items with a longer list (they lead the queue)
    CU(cudaMemcpyAsync(&((TileStats *)ctx->stats.p)->n_big, (const uint32_t *)ctx->hist.p + WIN_CAP, sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    // the queue holds one item per cell plus the extra slices of split cells: size it from the counts k1_hist produced
    CU(cudaMemcpyAsync(&ts, ctx->stats.p, sizeof(ts), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->n_queue = (uint32_t)ts.n_items;
    ctx->n_split = (uint32_t)ts.n_split;
    ctx->n_big = (uint32_t)ts.n_big;
    if (int rc = reserve(ctx, ctx->queue, sizeof(uint32_t) * (size_t)std::max<uint32_t>(ctx->n_queue, 1))) return rc;
    if (int rc = reserve(ctx, ctx->cell_slot, sizeof(uint32_t) * (size_t)G.ncx * G.ncy)) return rc;
    CU(cudaMemsetAsync((uint32_t *)ctx->counter.p + 2, 0, sizeof(uint32_t), ctx->stream));
    k1_order<<<gb, tb, 0, ctx->stream>>>(loff, (uint32_t *)ctx->hist.p, (uint32_t *)ctx->queue.p, (uint32_t *)ctx->cell_slot.p, (uint32_t *)ctx->counter.p + 2, cyb, cye);
    ctx->launches += 5;
    CU(cudaGetLastError());
    return 0;
}

template <int ERF, int Q, bool PACK, int MINB, bool CONTIG, bool WIN = false>
void launch_k2c(vrt_cuda_ctx *ctx, const RenderArgs &a)
{
    int per_sm = 1;
    constexpr int CTA_WARPS = k2_cta_warps(Q, MINB);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2_render<ERF, Q, PACK, MINB, CONTIG, WIN>, CTA_WARPS * 32, 0);
    if (per_sm < 1) per_sm = 1;
    const uint32_t want = (a.n_queue + CTA_WARPS - 1) / CTA_WARPS;
    const uint32_t grid = std::max(1u, std::min(want, (uint32_t)(ctx->sm_count * per_sm)));
    k2_render<ERF, Q, PACK, MINB, CONTIG, WIN><<<grid, CTA_WARPS * 32, 0, ctx->stream>>>(a);
}

template <int ERF, int Q, bool PACK, int MINB = (Q <= 4 ? 2 : 1)>
void launch_k2(vrt_cuda_ctx *ctx, const RenderArgs &a)
{
    if (a.list_idx == nullptr) launch_k2c<ERF, Q, PACK, MINB, true>(ctx, a);
    else launch_k2c<ERF, Q, PACK, MINB, false>(ctx, a);
}

template <int ERF>
int dispatch_k2(vrt_cuda_ctx *ctx, const RenderArgs &a)
{
    // Q = 8 (236 registers, 8 warps/SM) measured fastest on B200: instruction-level parallelism across 40 independent
    // sample chains per thread beats occupancy (tools/tune_k2.py); short lists waste less padding with Q = 4
    const int q = ctx->tune_q ? ctx->tune_q : ((ctx->n_lists && ctx->n_entries / ctx->n_lists >= 24) ? 8 : 4);
    const bool p = ctx->tune_pack != 0;
    if (a.window)
    {
        // depth-window mode (depth-sorted index lists).  The queue is in descending list length, so the cells whose list
        // does not fit the per-warp cache of k2_window lead it: they go to k2_render's in-loop saturation test, the rest
        // to k2_window.
        const uint32_t n_big = std::min(ctx->n_big, a.n_queue);
        if (n_big)
        {
            RenderArgs big = a;
            big.n_queue = n_big;
            launch_k2c<ERF, 8, true, 1, false, true>(ctx, big);
            ctx->render_launches++;
        }
        if (a.n_queue > n_big)
        {
            const size_t smem = sizeof(WinSmem) * K2_WARPS;
            if (!ctx->win_attr[ERF]) // per device: the opt-in for > 48 KB of dynamic shared memory
            {
                cudaFuncSetAttribute(k2_window<ERF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                ctx->win_attr[ERF] = true;
            }
            const uint32_t want = (a.n_queue - n_big + K2_WARPS - 1) / K2_WARPS;
            const uint32_t grid = std::max(1u, std::min(want, (uint32_t)ctx->sm_count));
            k2_window<ERF><<<grid, K2_WARPS * 32, smem, ctx->stream>>>(a, n_big);
        }
        return 0;
    }
    // Variants that were measured and dropped (tools/tune_k2.py): Q = 2, 6, 10; 3-4 CTAs/SM by register cap;
    // two occluders per step; sign taken before the reciprocal.  Kept: Q = 8 and 4, packed, and scalar math for comparison.
    // Q = 8 in 4-warp CTAs, three per SM (168 registers, 12 warps/SM): fastest on the depth-sorted bounded lists, where most
    // occluders take the cheap sign-uniform body and more warps are needed to cover the per-occluder setup
    if (ctx->tune_pack == 2 || (ctx->tune_q == 0 && ctx->tune_pack == 1 && q == 8 && ctx->lists_sorted))
    {
        launch_k2<ERF, 8, true, 3>(ctx, a);
        return 0;
    }
    switch (q)
    {
    case 4: p ? launch_k2<ERF, 4, true>(ctx, a) : launch_k2<ERF, 4, false>(ctx, a); break;
    case 8: p ? launch_k2<ERF, 8, true>(ctx, a) : launch_k2<ERF, 8, false>(ctx, a); break;
    default: return fail(ctx, VRT_CUDA_E_INVALID, "unsupported emitter block %d", q);
    }
    return 0;
}

template <int ERFV, int EXPV>
void launch_variant(vrt_cuda_ctx *ctx, const RenderArgs &a)
{
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2_variant<ERFV, EXPV>, K2_WARPS * 32, 0);
    if (per_sm < 1) per_sm = 1;
    const uint32_t want = (a.n_queue + K2_WARPS - 1) / K2_WARPS;
    const uint32_t grid = std::max(1u, std::min(want, (uint32_t)(ctx->sm_count * per_sm)));
    k2_variant<ERFV, EXPV><<<grid, K2_WARPS * 32, 0, ctx->stream>>>(a);
}

template <int EXPV>
int dispatch_variant_erf(vrt_cuda_ctx *ctx, const RenderArgs &a, int erfv)
{
    switch (erfv)
    {
    case ERFV_AS: launch_variant<ERFV_AS, EXPV>(ctx, a); break;
    case ERFV_EXACT: launch_variant<ERFV_EXACT, EXPV>(ctx, a); break;
    case ERFV_SPLINE: launch_variant<ERFV_SPLINE, EXPV>(ctx, a); break;
    case ERFV_SPLINE_MIRROR: launch_variant<ERFV_SPLINE_MIRROR, EXPV>(ctx, a); break;
    case ERFV_TAYLOR: launch_variant<ERFV_TAYLOR, EXPV>(ctx, a); break;
    default: return fail(ctx, VRT_CUDA_E_INVALID, "unknown erf approximation %d", erfv);
    }
    return 0;
}

int dispatch_variant(vrt_cuda_ctx *ctx, const RenderArgs &a, int erfv, int expv)
{
    switch (expv)
    {
    case EXPV_EXACT: return dispatch_variant_erf<EXPV_EXACT>(ctx, a, erfv);
    case EXPV_FAST: return dispatch_variant_erf<EXPV_FAST>(ctx, a, erfv);
    case EXPV_SPLINE: return dispatch_variant_erf<EXPV_SPLINE>(ctx, a, erfv);
    default: return fail(ctx, VRT_CUDA_E_INVALID, "unknown exp approximation %d", expv);
    }
}

int upload_approx_tables(vrt_cuda_ctx *ctx)
{
    ApproxTables t{};
    for (int i = 0; i < VRT_SPLINE_ERF_SEGMENTS; ++i)
        t.erf_coef[i] = make_float4(vrt_spline_erf_coef[i][0], vrt_spline_erf_coef[i][1], vrt_spline_erf_coef[i][2], vrt_spline_erf_coef[i][3]);
    for (int i = 0; i < VRT_SPLINE_EXP_SEGMENTS; ++i)
        t.exp_coef[i] = make_float4(vrt_spline_exp_coef[i][0], vrt_spline_exp_coef[i][1], vrt_spline_exp_coef[i][2], vrt_spline_exp_coef[i][3]);
    std::memcpy(t.erf_knot, vrt_spline_erf_knot, sizeof(t.erf_knot));
    std::memcpy(t.exp_knot, vrt_spline_exp_knot, sizeof(t.exp_knot));
    CU(cudaMemcpyToSymbol(c_approx, &t, sizeof(t)));
    return 0;
}
} // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C"
{
void vrt_cuda_destroy(vrt_cuda_ctx *ctx);
int vrt_cuda_abi_version(void) { return VRT_CUDA_ABI_VERSION; }

int vrt_cuda_create(int device, vrt_cuda_ctx **ctx_out)
{
    vrt_cuda_ctx *ctx = nullptr;
    if (!ctx_out) return fail(ctx, VRT_CUDA_E_INVALID, "ctx_out is NULL");
    *ctx_out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(ctx, VRT_CUDA_E_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(ctx, VRT_CUDA_E_INVALID, "device %d out of range (0..%d)", device, count - 1);
    cudaDeviceProp prop;
    CU(cudaSetDevice(device));
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(ctx, VRT_CUDA_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    vrt_cuda_ctx *c = new vrt_cuda_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    ctx = c;
    cudaError_t e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e2 == cudaSuccess; ++i) e2 = cudaEventCreate(&c->ev[i]);
    if (e2 != cudaSuccess)
    {
        g_create_error = std::string("stream/event creation failed: ") + cudaGetErrorString(e2);
        delete c;
        return VRT_CUDA_E_CUDA;
    }
    if (upload_approx_tables(c) != 0)
    {
        g_create_error = c->err;
        vrt_cuda_destroy(c);
        return VRT_CUDA_E_CUDA;
    }
    *ctx_out = c;
    return 0;
}

void vrt_cuda_destroy(vrt_cuda_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->aos, &ctx->rec, &ctx->cullrec, &ctx->ccounts, &ctx->coffsets, &ctx->cidx,
                      &ctx->hist, &ctx->queue, &ctx->stats, &ctx->counter, &ctx->rowcost, &ctx->scan_tmp, &ctx->cell_slot, &ctx->partial, &ctx->out_image, &ctx->out_rad, &ctx->tile_aos, &ctx->tile_off, &ctx->lit_offsets, &ctx->lit_idx};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    for (int i = 0; i < 4; ++i)
        for (DevBuf *b : {&ctx->lvl_counts[i], &ctx->lvl_offsets[i], &ctx->lvl_idx[i]})
            if (b->p) cudaFree(b->p);
    if (ctx->lvl_group_off.p) cudaFree(ctx->lvl_group_off.p);
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *vrt_cuda_last_error(const vrt_cuda_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
int vrt_cuda_device(const vrt_cuda_ctx *ctx) { return ctx ? ctx->device : -1; }
uint64_t vrt_cuda_stream(vrt_cuda_ctx *ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->stream : 0; }

int vrt_cuda_approx_table(vrt_cuda_ctx *ctx, int fn, const float *x, float *y, uint64_t n)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (fn < 0 || fn > VRT_CUDA_FN_SPLINE_EXP) return fail(ctx, VRT_CUDA_E_INVALID, "unknown function id %d", fn);
    if (n == 0) return 0;
    if (!x || !y) return fail(ctx, VRT_CUDA_E_INVALID, "x / y is NULL");
    CU(cudaSetDevice(ctx->device));
    if (int rc = reserve(ctx, ctx->out_rad, 2 * n * sizeof(float))) return rc;
    float *dx = (float *)ctx->out_rad.p, *dy = dx + n;
    CU(cudaMemcpyAsync(dx, x, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    k_approx_table<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(fn, dx, dy, n);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(y, dy, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int vrt_cuda_sync(vrt_cuda_ctx *ctx)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int set_gaussians_impl(vrt_cuda_ctx *ctx, const float *aos, uint64_t n, cudaMemcpyKind kind)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (n && !aos) return fail(ctx, VRT_CUDA_E_INVALID, "aos is NULL");
    if (n > 0xFFFFFFF0ull) return fail(ctx, VRT_CUDA_E_INVALID, "too many Gaussians");
    CU(cudaSetDevice(ctx->device));
    if (int rc = reserve(ctx, ctx->aos, std::max<size_t>(n, 1) * 40)) return rc;
    if (n) CU(cudaMemcpyAsync(ctx->aos.p, aos, n * 40, kind, ctx->stream));
    ctx->n_gauss = n;
    ctx->have_lists = false;
    return 0;
}

int vrt_cuda_set_gaussians(vrt_cuda_ctx *ctx, const float *aos, uint64_t n) { return set_gaussians_impl(ctx, aos, n, cudaMemcpyHostToDevice); }
int vrt_cuda_set_gaussians_device(vrt_cuda_ctx *ctx, const float *aos_dev, uint64_t n) { return set_gaussians_impl(ctx, aos_dev, n, cudaMemcpyDeviceToDevice); }

// Internal tuning knob (emitter block size Q in {2,4,6,8}, packed f32x2 math on/off); used by the benchmarks.
int vrt_cuda_set_tuning(vrt_cuda_ctx *ctx, int q, int pack)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (q != 0 && q != 4 && q != 8) return fail(ctx, VRT_CUDA_E_INVALID, "Q must be 0 (auto), 4 or 8");
    ctx->tune_q = q;
    ctx->tune_pack = pack;
    return 0;
}

// Measures the FP32 FMA throughput of the device (TFLOP/s, FMA = 2 flops) with scalar FFMA (packed = 0) or FFMA2.
int vrt_cuda_fp32_peak(vrt_cuda_ctx *ctx, int packed, double *tflops_out)
{
    if (!ctx || !tflops_out) return VRT_CUDA_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (int rc = reserve(ctx, ctx->counter, sizeof(uint32_t) * 4)) return rc;
    const int iters = 4096, blocks = ctx->sm_count * 8, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep)
    {
        CU(cudaEventRecord(ctx->ev[2], ctx->stream));
        if (packed) k_fp32_peak<true><<<blocks, threads, 0, ctx->stream>>>((float *)ctx->counter.p, iters, 0.999f, 0.001f);
        else k_fp32_peak<false><<<blocks, threads, 0, ctx->stream>>>((float *)ctx->counter.p, iters, 0.999f, 0.001f);
        CU(cudaEventRecord(ctx->ev[3], ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
        if (rep > 0 && ms < best) best = ms;
    }
    *tflops_out = (double)blocks * threads * 16.0 * iters * 2.0 / (best * 1e-3) / 1e12;
    return 0;
}

// Measures the ceiling of K2's inner-term instruction mix (terms/s) with `pairs` independent FFMA2 pairs per thread.
int vrt_cuda_term_peak(vrt_cuda_ctx *ctx, int pairs, int ctas_per_sm, double *terms_per_s_out)
{
    if (!ctx || !terms_per_s_out) return VRT_CUDA_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (int rc = reserve(ctx, ctx->counter, sizeof(uint32_t) * 4)) return rc;
    const int iters = 2048, blocks = ctx->sm_count * (ctas_per_sm > 0 ? ctas_per_sm : 1), threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep)
    {
        CU(cudaEventRecord(ctx->ev[2], ctx->stream));
        switch (pairs)
        {
        case 5: k_term_peak<5, false><<<blocks, threads, 0, ctx->stream>>>((float *)ctx->counter.p, iters, 1.1f, -0.3f, 0.5f); break;
        case 10: k_term_peak<10, false><<<blocks, threads, 0, ctx->stream>>>((float *)ctx->counter.p, iters, 1.1f, -0.3f, 0.5f); break;
        case 20: k_term_peak<20, false><<<blocks, threads, 0, ctx->stream>>>((float *)ctx->counter.p, iters, 1.1f, -0.3f, 0.5f); break;
        case -10: k_term_peak<10, true><<<blocks, threads, 0, ctx->stream>>>((float *)ctx->counter.p, iters, 1.1f, -0.3f, 0.5f); break; // sign-uniform body
        case -20: k_term_peak<20, true><<<blocks, threads, 0, ctx->stream>>>((float *)ctx->counter.p, iters, 1.1f, -0.3f, 0.5f); break;
        default: return fail(ctx, VRT_CUDA_E_INVALID, "pairs must be 5, 10, 20 (signed body) or -10, -20 (sign-uniform body)");
        }
        CU(cudaEventRecord(ctx->ev[3], ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
        if (rep > 0 && ms < best) best = ms;
    }
    *terms_per_s_out = (double)blocks * threads * 2.0 * std::abs(pairs) * iters / (best * 1e-3);
    return 0;
}

// Pipe-mix probe: steps/s of (nf packed FMAs, nm MUFU.RCP, nl LOP3) per float2 chain step; see tools/probe_peaks.py.
int vrt_cuda_mix_peak(vrt_cuda_ctx *ctx, int nf, int nm, int nl, double *steps_per_s_out)
{
    if (!ctx || !steps_per_s_out) return VRT_CUDA_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (int rc = reserve(ctx, ctx->counter, sizeof(uint32_t) * 4)) return rc;
    const int iters = 1024, blocks = ctx->sm_count * 4, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep)
    {
        CU(cudaEventRecord(ctx->ev[2], ctx->stream));
        float *o = (float *)ctx->counter.p;
        const int key = nf * 100 + nm * 10 + nl;
        switch (key)
        {
        case 900: k_mix_peak<9, 0, 0><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 0.001f); break;
        case 902: k_mix_peak<9, 0, 2><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 0.001f); break;
        case 910: k_mix_peak<9, 1, 0><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 0.001f); break;
        case 920: k_mix_peak<9, 2, 0><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 0.001f); break;
        case 922: k_mix_peak<9, 2, 2><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 0.001f); break;
        case 20: k_mix_peak<0, 2, 0><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 0.001f); break;
        case 120: k_mix_peak<1, 2, 0><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 0.001f); break;
        case 420: k_mix_peak<4, 2, 0><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 0.001f); break;
        case 1820: k_mix_peak<18, 2, 0><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 0.001f); break;
        default: return fail(ctx, VRT_CUDA_E_INVALID, "unsupported mix %d", key);
        }
        CU(cudaEventRecord(ctx->ev[3], ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
        if (rep > 0 && ms < best) best = ms;
    }
    *steps_per_s_out = (double)blocks * threads * 16.0 * iters / (best * 1e-3);
    return 0;
}

// K0 + K1 for one frame: records, lists of the frame's list mode, depth sort, work queue
static int tile_build(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    FrameGeom G;
    std::vector<float> cxs, cys;
    if (int rc = make_geom(ctx, frame, -1, G, cxs, cys)) return rc;
    ctx->have_lists = false;
    ctx->lists_from_host = false;
    ctx->literal = false;
    ctx->geom = G;
    ctx->launches = 0;
    const uint64_t N = ctx->n_gauss;
    if (int rc = upload_geom(ctx, G, cxs, cys)) return rc;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));

    if (int rc = reserve(ctx, ctx->rec, sizeof(Rec) * std::max<uint64_t>(N, 1))) return rc;
    if (int rc = reserve(ctx, ctx->cullrec, 2 * sizeof(float4) * std::max<uint64_t>(N, 1))) return rc;
    if (N)
    {
        k0_prepare<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>((const float *)ctx->aos.p, N, (Rec *)ctx->rec.p, (float4 *)ctx->cullrec.p);
        ctx->launches++;
    }

    if (G.list_kind == 2)
    {
        // single list: all Gaussians, contiguous
        if (int rc = reserve(ctx, ctx->coffsets, sizeof(uint32_t) * 2)) return rc;
        const uint32_t off[2] = {0u, (uint32_t)N};
        CU(cudaMemcpyAsync(ctx->coffsets.p, off, sizeof(off), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream)); // off[] is on the stack
        ctx->n_lists = 1;
        ctx->n_entries = N;
    }
    else if (G.list_kind == 1)
    {
        const uint32_t nt = (uint32_t)(G.tiles_x * G.tiles_y);
        if (int rc = reserve(ctx, ctx->ccounts, sizeof(uint32_t) * nt)) return rc;
        if (int rc = reserve(ctx, ctx->coffsets, sizeof(uint32_t) * (nt + 1))) return rc;
        k1_cull_tiles<false><<<nt, 256, 0, ctx->stream>>>((const Rec *)ctx->rec.p, (const float4 *)ctx->cullrec.p, (uint32_t)N, (uint32_t *)ctx->ccounts.p, nullptr, nullptr, nt);
        k1_scan<<<1, 1024, 0, ctx->stream>>>((const uint32_t *)ctx->ccounts.p, (uint32_t *)ctx->coffsets.p, nt);
        uint32_t total = 0;
        CU(cudaMemcpyAsync(&total, (const uint32_t *)ctx->coffsets.p + nt, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (int rc = reserve(ctx, ctx->cidx, sizeof(uint32_t) * std::max<uint32_t>(total, 1))) return rc;
        k1_cull_tiles<true><<<nt, 256, 0, ctx->stream>>>((const Rec *)ctx->rec.p, (const float4 *)ctx->cullrec.p, (uint32_t)N, nullptr, (const uint32_t *)ctx->coffsets.p, (uint32_t *)ctx->cidx.p, nt);
        ctx->launches += 3;
        ctx->n_lists = nt;
        ctx->n_entries = total;
    }
    else
    {
        // levels from coarse to fine: groups of 64x128, 16x32, 4x8 and 1x1 cells (512, 128, 32 pixels square, then the
        // 8x4-pixel cell); a level is dropped when it would not be coarser than the whole grid
        std::vector<CullLevel> levels;
        const int gxs[4] = {64, 16, 4, 1}, gys[4] = {128, 32, 8, 1};
        for (int i = 0; i < 4; ++i)
        {
            if (i < 3 && gxs[i] >= G.ncx && gys[i] >= G.ncy && !levels.empty()) continue;
            if (i < 3 && gxs[i + 1] >= G.ncx && gys[i + 1] >= G.ncy) continue; // the next finer level is still one group
            CullLevel L{};
            L.gx = gxs[i]; L.gy = gys[i];
            L.ngx = (G.ncx + L.gx - 1) / L.gx; L.ngy = (G.ncy + L.gy - 1) / L.gy;
            L.is_root = levels.empty() ? 1 : 0;
            L.n_seg = L.is_root ? (int)std::max<uint64_t>(1, (N + ROOT_SEG - 1) / ROOT_SEG) : 1;
            if (!levels.empty()) { L.pgx = levels.back().gx; L.pgy = levels.back().gy; L.pngx = levels.back().ngx; }
            levels.push_back(L);
        }
        const uint32_t *parent_off = nullptr, *parent_idx = nullptr;
        for (size_t li = 0; li < levels.size(); ++li)
        {
            const CullLevel &L = levels[li];
            const bool last = li + 1 == levels.size();
            const uint64_t groups = (uint64_t)L.ngx * L.ngy, work = groups * L.n_seg;
            if (work > 0x7FFFFFFFull / 32) return fail(ctx, VRT_CUDA_E_INVALID, "scene x image too large for culling level %zu", li);
            DevBuf &cnt = last ? ctx->ccounts : ctx->lvl_counts[li];
            DevBuf &off = last ? ctx->coffsets : ctx->lvl_offsets[li];
            DevBuf &idx = last ? ctx->cidx : ctx->lvl_idx[li];
            if (int rc = reserve(ctx, cnt, sizeof(uint32_t) * work)) return rc;
            if (int rc = reserve(ctx, off, sizeof(uint32_t) * (work + 1))) return rc;
            const unsigned grid = (unsigned)((work * 32 + 255) / 256);
            k1_cull<false><<<grid, 256, 0, ctx->stream>>>((const Rec *)ctx->rec.p, (const float4 *)ctx->cullrec.p, (uint32_t)N, L, parent_off, parent_idx,
                                                         (uint32_t *)cnt.p, nullptr, nullptr, (uint32_t)work);
            if (int rc = scan_u32(ctx, (const uint32_t *)cnt.p, (uint32_t *)off.p, (uint32_t)work)) return rc;
            uint32_t total = 0;
            CU(cudaMemcpyAsync(&total, (const uint32_t *)off.p + work, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            if (int rc = reserve(ctx, idx, sizeof(uint32_t) * std::max<uint32_t>(total, 1))) return rc;
            k1_cull<true><<<grid, 256, 0, ctx->stream>>>((const Rec *)ctx->rec.p, (const float4 *)ctx->cullrec.p, (uint32_t)N, L, parent_off, parent_idx, nullptr,
                                                        (const uint32_t *)off.p, (uint32_t *)idx.p, (uint32_t)work);
            ctx->launches += 2;
            parent_idx = (const uint32_t *)idx.p;
            if (L.n_seg > 1)
            {
                // per-(group, segment) offsets -> per-group offsets for the next level
                if (int rc = reserve(ctx, ctx->lvl_group_off, sizeof(uint32_t) * (groups + 1))) return rc;
                k1_group_offsets<<<(unsigned)((groups + 256) / 256), 256, 0, ctx->stream>>>((const uint32_t *)off.p, (uint32_t *)ctx->lvl_group_off.p, (uint32_t)groups, L.n_seg);
                ctx->launches++;
                parent_off = (const uint32_t *)ctx->lvl_group_off.p;
                if (last)
                {
                    // a single-level hierarchy with a segmented root: the cell offsets are the group offsets
                    CU(cudaMemcpyAsync(ctx->coffsets.p, ctx->lvl_group_off.p, sizeof(uint32_t) * (groups + 1), cudaMemcpyDeviceToDevice, ctx->stream));
                }
            }
            else parent_off = (const uint32_t *)off.p;
            if (last)
            {
                ctx->n_lists = (uint32_t)groups;
                ctx->n_entries = total;
            }
        }
    }
    ctx->lists_sorted = false;
    // bounded per-cell lists are always depth-sorted: the order is deterministic (ties by index) and makes most occluders
    // sign-uniform for an emitter block (K2), and saturated in depth-window mode
    if (G.list_kind == 0 && ctx->n_entries)
    {
        k1_sort_cells<<<(ctx->n_lists + 3) / 4, 128, 0, ctx->stream>>>((const float4 *)ctx->cullrec.p, (const uint32_t *)ctx->coffsets.p, (uint32_t *)ctx->cidx.p, ctx->n_lists);
        ctx->launches++;
        ctx->lists_sorted = true;
    }
    CU(cudaGetLastError());
    if (int rc = build_queue(ctx)) return rc;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaEventElapsedTime(&ctx->ms_tile, ctx->ev[0], ctx->ev[1]));
    ctx->have_lists = true;
    return 0;
}

// Literal list modes (the reference's tile membership, the ALL list, caller-supplied tiles_t) name far more Gaussians per
// pixel than the pixel can see: an entry farther than VISIBLE_SIGMAS sigma from every ray of a cell has a weight of exactly
// 0 there and adds exactly 0 to every sum.  Instead of discovering that per (occluder, emitter block) inside the n^2 loop,
// K1 intersects the literal lists with "visible from the cell" once per frame and K2 walks those per-cell lists (depth-
// sorted, split, queued like the bounded modes).  The literal lists stay available to vrt_cuda_get_lists and the literal
// counts to vrt_cuda_stats; the image is the literal one up to the order of the fp32 sums.  VRT_CUDA_NO_SKIP turns this off.
static int keep_literal(vrt_cuda_ctx *ctx)
{
    CU(cudaMemcpyAsync(&ctx->lit_stats, ctx->stats.p, sizeof(TileStats), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::swap(ctx->coffsets, ctx->lit_offsets);
    std::swap(ctx->cidx, ctx->lit_idx);
    ctx->lit_kind = ctx->geom.list_kind;
    ctx->lit_n_lists = ctx->n_lists;
    ctx->lit_n_entries = ctx->n_entries;
    return 0;
}

int vrt_cuda_tile(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame)
{
    if (int rc = tile_build(ctx, frame)) return rc;
    const uint32_t lm = frame->flags & VRT_CUDA_LIST_MASK;
    if ((lm != VRT_CUDA_LIST_REFERENCE && lm != VRT_CUDA_LIST_ALL) || (frame->flags & VRT_CUDA_NO_SKIP) || ctx->n_gauss == 0) return 0;
    ctx->have_lists = false;
    if (int rc = keep_literal(ctx)) return rc;
    const float ms_literal = ctx->ms_tile;
    const uint32_t launches_literal = ctx->launches;
    vrt_cuda_frame visible = *frame;
    visible.flags = (frame->flags & ~VRT_CUDA_LIST_MASK) | (lm == VRT_CUDA_LIST_REFERENCE ? VRT_CUDA_LIST_REFERENCE_BOUND : VRT_CUDA_LIST_BOUND);
    visible.bound_sigmas = VISIBLE_SIGMAS;
    if (int rc = tile_build(ctx, &visible)) return rc;
    ctx->ms_tile += ms_literal;
    ctx->launches += launches_literal;
    ctx->literal = true;
    return 0;
}

int vrt_cuda_set_tile_lists(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame, const float *aos_concat, const uint64_t *offsets, uint64_t n_tiles)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (!offsets) return fail(ctx, VRT_CUDA_E_INVALID, "offsets is NULL");
    FrameGeom G;
    std::vector<float> cxs, cys;
    if (int rc = make_geom(ctx, frame, 1, G, cxs, cys)) return rc;
    if (n_tiles != (uint64_t)G.tiles_x * G.tiles_y) return fail(ctx, VRT_CUDA_E_INVALID, "n_tiles %llu != tiles_x*tiles_y %d", (unsigned long long)n_tiles, G.tiles_x * G.tiles_y);
    const uint64_t total = offsets[n_tiles];
    if (total > 0xFFFFFFF0ull) return fail(ctx, VRT_CUDA_E_INVALID, "tile lists too long");
    if (total && !aos_concat) return fail(ctx, VRT_CUDA_E_INVALID, "aos_concat is NULL");
    std::vector<uint32_t> off32(n_tiles + 1);
    for (uint64_t t = 0; t <= n_tiles; ++t)
    {
        if (t && offsets[t] < offsets[t - 1]) return fail(ctx, VRT_CUDA_E_INVALID, "offsets must be non-decreasing");
        off32[t] = (uint32_t)offsets[t];
    }
    ctx->have_lists = false;
    ctx->literal = false;
    ctx->geom = G;
    ctx->launches = 0;
    if (int rc = upload_geom(ctx, G, cxs, cys)) return rc;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (int rc = reserve(ctx, ctx->tile_aos, std::max<uint64_t>(total, 1) * 40)) return rc;
    if (int rc = reserve(ctx, ctx->rec, sizeof(Rec) * std::max<uint64_t>(total, 1))) return rc;
    if (int rc = reserve(ctx, ctx->cullrec, 2 * sizeof(float4) * std::max<uint64_t>(total, 1))) return rc;
    if (int rc = reserve(ctx, ctx->coffsets, sizeof(uint32_t) * (n_tiles + 1))) return rc;
    if (total) CU(cudaMemcpyAsync(ctx->tile_aos.p, aos_concat, total * 40, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->coffsets.p, off32.data(), sizeof(uint32_t) * (n_tiles + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (total)
    {
        k0_prepare<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const float *)ctx->tile_aos.p, total, (Rec *)ctx->rec.p, (float4 *)ctx->cullrec.p);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    ctx->n_lists = (uint32_t)n_tiles;
    ctx->n_entries = total;
    ctx->lists_from_host = true;
    ctx->lists_sorted = false;
    if (int rc = build_queue(ctx)) return rc;
    if (total && !(frame->flags & VRT_CUDA_NO_SKIP))
    {
        // visible lists (see vrt_cuda_tile): one cull level, every 8x4 cell scans the record range of its own tile
        CU(cudaStreamSynchronize(ctx->stream)); // off32 (stack) is consumed before the offsets buffer changes hands
        if (int rc = keep_literal(ctx)) return rc;
        FrameGeom V = G;
        V.list_kind = 0;
        V.use_ref = 0;
        V.use_bound = 1;
        V.bound_k = VISIBLE_SIGMAS;
        ctx->geom = V;
        if (int rc = upload_geom(ctx, V, cxs, cys)) return rc;
        CullLevel L{};
        L.gx = L.gy = 1;
        L.ngx = V.ncx; L.ngy = V.ncy;
        L.pgx = V.cptx; L.pgy = V.cpty; L.pngx = V.tiles_x;
        L.is_root = 0; L.n_seg = 1;
        const uint64_t cells = (uint64_t)V.ncx * V.ncy;
        if (cells > 0x7FFFFFFFull / 32) return fail(ctx, VRT_CUDA_E_INVALID, "image too large");
        if (int rc = reserve(ctx, ctx->ccounts, sizeof(uint32_t) * cells)) return rc;
        if (int rc = reserve(ctx, ctx->coffsets, sizeof(uint32_t) * (cells + 1))) return rc;
        const unsigned grid = (unsigned)((cells * 32 + 255) / 256);
        k1_cull<false><<<grid, 256, 0, ctx->stream>>>((const Rec *)ctx->rec.p, (const float4 *)ctx->cullrec.p, (uint32_t)total, L, (const uint32_t *)ctx->lit_offsets.p, nullptr,
                                                     (uint32_t *)ctx->ccounts.p, nullptr, nullptr, (uint32_t)cells);
        if (int rc = scan_u32(ctx, (const uint32_t *)ctx->ccounts.p, (uint32_t *)ctx->coffsets.p, (uint32_t)cells)) return rc;
        uint32_t kept = 0;
        CU(cudaMemcpyAsync(&kept, (const uint32_t *)ctx->coffsets.p + cells, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (int rc = reserve(ctx, ctx->cidx, sizeof(uint32_t) * std::max<uint32_t>(kept, 1))) return rc;
        k1_cull<true><<<grid, 256, 0, ctx->stream>>>((const Rec *)ctx->rec.p, (const float4 *)ctx->cullrec.p, (uint32_t)total, L, (const uint32_t *)ctx->lit_offsets.p, nullptr, nullptr,
                                                    (const uint32_t *)ctx->coffsets.p, (uint32_t *)ctx->cidx.p, (uint32_t)cells);
        ctx->launches += 2;
        ctx->n_lists = (uint32_t)cells;
        ctx->n_entries = kept;
        if (kept)
        {
            k1_sort_cells<<<(ctx->n_lists + 3) / 4, 128, 0, ctx->stream>>>((const float4 *)ctx->cullrec.p, (const uint32_t *)ctx->coffsets.p, (uint32_t *)ctx->cidx.p, ctx->n_lists);
            ctx->launches++;
            ctx->lists_sorted = true;
        }
        CU(cudaGetLastError());
        if (int rc = build_queue(ctx)) return rc;
        ctx->literal = true;
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream)); // off32 lives on this stack frame
    CU(cudaEventElapsedTime(&ctx->ms_tile, ctx->ev[0], ctx->ev[1]));
    ctx->have_lists = true;
    return 0;
}

int vrt_cuda_get_lists(vrt_cuda_ctx *ctx, uint32_t *counts_out, uint64_t counts_cap, uint32_t *idx_out, uint64_t idx_cap, uint64_t *n_cells_out, uint64_t *n_entries_out)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (!ctx->have_lists) return fail(ctx, VRT_CUDA_E_STATE, "no lists: call vrt_cuda_tile first");
    CU(cudaSetDevice(ctx->device));
    // literal modes report the literal lists (the membership of tile_gaussians), not the visible subsets K2 walks
    const uint32_t n_lists = ctx->literal ? ctx->lit_n_lists : ctx->n_lists;
    const uint64_t n_entries = ctx->literal ? ctx->lit_n_entries : ctx->n_entries;
    const int kind = ctx->literal ? ctx->lit_kind : ctx->geom.list_kind;
    const DevBuf &offsets = ctx->literal ? ctx->lit_offsets : ctx->coffsets;
    const DevBuf &indices = ctx->literal ? ctx->lit_idx : ctx->cidx;
    if (n_cells_out) *n_cells_out = n_lists;
    if (n_entries_out) *n_entries_out = n_entries;
    if (counts_out)
    {
        if (counts_cap < n_lists) return fail(ctx, VRT_CUDA_E_INVALID, "counts_cap too small");
        std::vector<uint32_t> off(n_lists + 1);
        CU(cudaMemcpyAsync(off.data(), offsets.p, sizeof(uint32_t) * off.size(), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        for (uint32_t i = 0; i < n_lists; ++i) counts_out[i] = off[i + 1] - off[i];
    }
    if (idx_out)
    {
        if (idx_cap < n_entries) return fail(ctx, VRT_CUDA_E_INVALID, "idx_cap too small");
        if (kind == 2 || ctx->lists_from_host)
        {
            for (uint64_t i = 0; i < n_entries; ++i) idx_out[i] = (uint32_t)i;
        }
        else if (n_entries)
        {
            CU(cudaMemcpyAsync(idx_out, indices.p, sizeof(uint32_t) * n_entries, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        }
    }
    return 0;
}

int vrt_cuda_row_costs(vrt_cuda_ctx *ctx, double *rows_out, uint32_t rows_cap, uint32_t *n_rows_out, uint32_t *row_height_px_out)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (!ctx->have_lists) return fail(ctx, VRT_CUDA_E_STATE, "no lists: call vrt_cuda_tile first");
    CU(cudaSetDevice(ctx->device));
    const FrameGeom &G = ctx->geom;
    if (G.tile_h % CELL_H) return fail(ctx, VRT_CUDA_E_INVALID, "row costs need tile_h %% %d == 0", CELL_H);
    if (n_rows_out) *n_rows_out = (uint32_t)G.ncy;
    if (row_height_px_out) *row_height_px_out = CELL_H;
    if (rows_out)
    {
        if (rows_cap < (uint32_t)G.ncy) return fail(ctx, VRT_CUDA_E_INVALID, "rows_cap too small");
        CU(cudaMemcpyAsync(rows_out, ctx->rowcost.p, sizeof(double) * G.ncy, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

int vrt_cuda_render_device(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame, uint32_t *image_dev, float *radiance_dev, vrt_cuda_stats *stats)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (!ctx->have_lists) return fail(ctx, VRT_CUDA_E_STATE, "no lists: call vrt_cuda_tile or vrt_cuda_set_tile_lists first");
    if (!frame) return fail(ctx, VRT_CUDA_E_INVALID, "frame is NULL");
    CU(cudaSetDevice(ctx->device));
    const FrameGeom &G = ctx->geom;
    if ((int)frame->width != G.W || (int)frame->height != G.H) return fail(ctx, VRT_CUDA_E_INVALID, "frame size differs from the tiled frame");
    if (std::memcmp(frame->view, G.view, sizeof(G.view)) != 0 || std::memcmp(frame->origin, G.origin, sizeof(float) * 3) != 0)
        return fail(ctx, VRT_CUDA_E_STATE, "camera changed since vrt_cuda_tile: lists are per frame");
    {
        const int rb = (frame->row_begin == 0 && frame->row_end == 0) ? 0 : (int)frame->row_begin;
        const int re = (frame->row_begin == 0 && frame->row_end == 0) ? G.H : (int)frame->row_end;
        if (rb != G.row_begin || re != G.row_end) return fail(ctx, VRT_CUDA_E_STATE, "row band differs from the tiled frame");
    }
    const uint32_t approx_erf = frame->flags & VRT_CUDA_APPROX_ERF_MASK, approx_exp = frame->flags & VRT_CUDA_APPROX_EXP_MASK;
    if ((approx_erf || approx_exp) && (frame->flags & VRT_CUDA_DEPTH_WINDOW))
        return fail(ctx, VRT_CUDA_E_INVALID, "the depth window needs a saturating odd erf: not available with VRT_CUDA_APPROX_*");
    if (approx_exp > VRT_CUDA_APPROX_EXP_SPLINE) return fail(ctx, VRT_CUDA_E_INVALID, "unknown exp approximation");
    RenderArgs a{};
    a.rec = (const Rec *)ctx->rec.p;
    a.list_off = (const uint32_t *)ctx->coffsets.p;
    a.list_idx = (G.list_kind == 2 || (ctx->lists_from_host && !ctx->literal)) ? nullptr : (const uint32_t *)ctx->cidx.p;
    if (ctx->literal && (frame->flags & VRT_CUDA_NO_SKIP))
        return fail(ctx, VRT_CUDA_E_STATE, "these lists were built with culling: pass VRT_CUDA_NO_SKIP to vrt_cuda_tile / vrt_cuda_set_tile_lists as well");
    a.queue = (const uint32_t *)ctx->queue.p;
    a.n_queue = ctx->n_queue;
    a.counter = (uint32_t *)ctx->counter.p;
    a.image = image_dev;
    a.radiance = (float4 *)radiance_dev;
    a.terms_exec = &((TileStats *)ctx->stats.p)->terms_exec;
    a.terms_sat = &((TileStats *)ctx->stats.p)->terms_sat;
    a.window = ((frame->flags & VRT_CUDA_DEPTH_WINDOW) && a.list_idx != nullptr) ? 1u : 0u;
    a.cell_slot = (const uint32_t *)ctx->cell_slot.p;
    if (ctx->n_split)
        if (int rc = reserve(ctx, ctx->partial, (size_t)ctx->n_split * 32 * sizeof(float4))) return rc;
    a.partial = (float4 *)ctx->partial.p;
    const bool bounded = G.use_bound != 0;
    a.skip_thresh = (frame->flags & VRT_CUDA_NO_SKIP) ? -1.f : ((bounded && !ctx->literal) ? std::exp2(-0.5f * G.bound_k * G.bound_k * LOG2E) : 0.f);
    a.quant_nearest = (frame->flags & VRT_CUDA_QUANT_NEAREST) ? 1u : 0u;
    a.alpha_from_w = (frame->flags & VRT_CUDA_ALPHA_FROM_W) ? 1u : 0u;
    CU(cudaMemsetAsync(ctx->counter.p, 0, sizeof(uint32_t) * 4, ctx->stream));
    CU(cudaMemsetAsync(&((TileStats *)ctx->stats.p)->terms_exec, 0, 2 * sizeof(unsigned long long), ctx->stream));
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    ctx->render_launches = 1;
    int rc;
    if (approx_erf || approx_exp)
    {
        // alternative approximations: an erf selection overrides the ERF_AS / ERF_EXACT bit
        const int erfv = approx_erf ? (int)(approx_erf >> 8) + 1 : (((frame->flags & VRT_CUDA_ERF_MASK) == VRT_CUDA_ERF_EXACT) ? ERFV_EXACT : ERFV_AS);
        rc = dispatch_variant(ctx, a, erfv, (int)(approx_exp >> 10));
    }
    else rc = ((frame->flags & VRT_CUDA_ERF_MASK) == VRT_CUDA_ERF_EXACT) ? dispatch_k2<1>(ctx, a) : dispatch_k2<0>(ctx, a);
    if (rc) return rc;
    if (ctx->n_split)
    {
        const uint32_t ncells = (uint32_t)((ctx->cy_end - ctx->cy_begin) * G.ncx);
        k3_combine<<<(unsigned)(((uint64_t)ncells * 32 + 255) / 256), 256, 0, ctx->stream>>>(a, (const uint32_t *)ctx->coffsets.p, ctx->cy_begin, ctx->cy_end);
        ctx->render_launches++;
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    if (stats)
    {
        TileStats ts;
        CU(cudaMemcpyAsync(&ts, ctx->stats.p, sizeof(ts), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        std::memset(stats, 0, sizeof(*stats));
        stats->n_gaussians = ctx->n_gauss;
        stats->n_cells = ctx->literal ? ctx->lit_n_lists : ctx->n_lists;
        stats->list_entries = ctx->literal ? ctx->lit_stats.entries : ts.entries;
        stats->max_list = (uint32_t)(ctx->literal ? ctx->lit_stats.max_list : ts.max_list);
        stats->n_launches = ctx->launches + ctx->render_launches;
        stats->terms_listed = ctx->literal ? ctx->lit_stats.terms_listed : ts.terms_listed;
        stats->terms_executed = (double)ts.terms_exec;
        stats->terms_saturated = (double)ts.terms_sat;
        stats->ms_tile = ctx->ms_tile;
        CU(cudaEventElapsedTime(&stats->ms_render, ctx->ev[2], ctx->ev[3]));
        stats->ms_total = stats->ms_tile + stats->ms_render;
    }
    return 0;
}

int vrt_cuda_render(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame, uint32_t *image, float *radiance, vrt_cuda_stats *stats)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (!frame) return fail(ctx, VRT_CUDA_E_INVALID, "frame is NULL");
    CU(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)frame->width * frame->height;
    if (image)
        if (int rc = reserve(ctx, ctx->out_image, npix * sizeof(uint32_t))) return rc;
    if (radiance)
        if (int rc = reserve(ctx, ctx->out_rad, npix * sizeof(float) * 4)) return rc;
    int rc = vrt_cuda_render_device(ctx, frame, image ? (uint32_t *)ctx->out_image.p : nullptr, radiance ? (float *)ctx->out_rad.p : nullptr, stats);
    if (rc) return rc;
    const FrameGeom &G = ctx->geom;
    const size_t row0 = (size_t)G.row_begin, rows = (size_t)(G.row_end - G.row_begin);
    if (image) CU(cudaMemcpyAsync(image + row0 * G.W, (uint32_t *)ctx->out_image.p + row0 * G.W, rows * G.W * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (radiance) CU(cudaMemcpyAsync(radiance + row0 * G.W * 4, (float *)ctx->out_rad.p + row0 * G.W * 4, rows * G.W * sizeof(float) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int vrt_cuda_frame_render(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame, uint32_t *image, float *radiance, vrt_cuda_stats *stats)
{
    if (int rc = vrt_cuda_tile(ctx, frame)) return rc;
    return vrt_cuda_render(ctx, frame, image, radiance, stats);
}
} // extern "C"
