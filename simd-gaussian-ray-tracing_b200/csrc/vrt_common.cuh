// vrt_common.cuh -- constants, the frame geometry (a kernel ARGUMENT: no per-device state), the per-Gaussian record and the
// erf device functions.  A fragment of vrt_cuda.cu: included there, in this order, inside its anonymous namespace.
// Not a stand-alone header.
#pragma once

constexpr int CELL_W = 8;          // pixels per cell (= one warp), x
constexpr int CELL_H = 4;          // y
constexpr int ROOT_SEG = 4096;     // Gaussians per root segment in the first cull level
constexpr int K2_WARPS = 8;        // warps per render CTA
// CTA shape per variant: Q = 8 with a 3-CTA/SM target uses 4-warp CTAs (register cap 168, 12 warps/SM)
__host__ __device__ constexpr int k2_cta_warps(int q, int minb) { return (q == 8 && minb == 3) ? 4 : K2_WARPS; }
constexpr int STAGE = 32;          // records staged per warp per step (one per lane)
constexpr int LONG_CAP_KEY = 832;  // = LONG_CAP (k2_long.cuh): longest list k2_band_long caches per CTA
constexpr int WIN_CAP = 152;       // longest list the depth-window kernel caches per warp ((WIN_CAP+1) * 128 B of prefix sums)
// Heavy cells (work ~ n^2) are split by emitter range into independent work items so one warp never owns a whole long list:
// a cell with more than 3 x slice entries becomes ceil(n / slice) items; the partial radiances are summed in slice order.
constexpr int SLICE_MAX = 256;     // emitters per item of a split cell: 256 on big frames (every item repeats pass A over the whole list:
                                   //   64 cost the 4M-Gaussian frame of profiles/r02_long_lists_ab.md 26 %), down to 8 when a frame has too few
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
    // float-accumulated tile centres (src/vrt/rt.cpp:47-49): 1024 + 1024 floats in the context's own device buffer
    const float *tile_cx, *tile_cy;
    const int *slice_dev; // the slice size chosen on the device during the tile call (read when `slice` is 0)
};
// The geometry travels by value in every kernel's parameter block (the frame is stateless on the device: two contexts, or
// two frames of one context, can be in flight on the same GPU -- the re-entrancy of the reference's entries).

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
