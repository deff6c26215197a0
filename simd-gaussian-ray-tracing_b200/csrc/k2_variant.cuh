// k2_variant.cuh -- K2'': the reference's alternative erf / exp approximations as device functions and the kernel that renders with them.
// A fragment of vrt_cuda.cu: included there, in this order, inside its anonymous namespace.  Not a stand-alone header.
#pragma once

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
    const FrameGeom &G = args.geom;
    const int lane = threadIdx.x & 31;
    const int lx = lane & (CELL_W - 1), ly = lane >> 3;
    constexpr float LN2 = 1.f / LOG2E;
    for (;;)
    {
        if (args.abort_flag && *args.abort_flag) break; // interrupted render
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(args.counter, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= args.n_queue) break;
        const uint32_t item = args.queue[qi];
        const uint32_t cell = item & ((1u << ITEM_CELL_BITS) - 1u), slice = item >> ITEM_CELL_BITS;
        const int cx = cell % G.ncx, cy = cell / G.ncx;
        int x0, y0, cw, ch;
        cell_rect(G, cx, cy, x0, y0, cw, ch);
        const int px = x0 + min(lx, cw - 1), py = y0 + min(ly, ch - 1);
        const bool live = lx < cw && ly < ch && py >= G.row_begin && py < G.row_end;
        const uint32_t n_live = __popc(__ballot_sync(0xffffffffu, live));
        const PixelRay ray = pixel_ray(G, px, py);
        const uint32_t lid = cell_list_id(G, cx, cy);
        const uint32_t off = args.list_off[lid];
        const uint32_t n = args.list_cnt[lid];
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
        const uint32_t q_begin = slot != NO_SLOT ? slice * frame_slice(G) : 0u, q_end = slot != NO_SLOT ? min(n, q_begin + frame_slice(G)) : n;
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
        else store_cell(args, G, px, py, live, Lr, Lg, Lb, La);
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
