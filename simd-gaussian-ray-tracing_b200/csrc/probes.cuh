// probes.cuh -- roofline probes: FP32 FMA peak, inner-term ceiling, pipe-mix chains.
// A fragment of vrt_cuda.cu: included there, in this order, inside its anonymous namespace.  Not a stand-alone header.
#pragma once

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
// Throughput of the approximation functions (the GPU counterpart of tests/approx_cycles.cpp, which counts CPU cycles per
// value with rdpmc): 8 independent argument chains per thread, each value nudges its own next argument so that nothing
// can be hoisted, arguments stay in the function's interesting range.
// ------------------------------------------------------------------------------------------------
template <int FN>
__device__ __forceinline__ float approx_fn(const ApproxTables &T, float x)
{
    switch (FN)
    {
    case VRT_CUDA_FN_SPLINE_ERF: return erf_approx<ERFV_SPLINE>(T, x);
    case VRT_CUDA_FN_SPLINE_ERF_MIRROR: return erf_approx<ERFV_SPLINE_MIRROR>(T, x);
    case VRT_CUDA_FN_TAYLOR_ERF: return erf_approx<ERFV_TAYLOR>(T, x);
    case VRT_CUDA_FN_AS_ERF: return erf_approx<ERFV_AS>(T, x);
    case VRT_CUDA_FN_ERF: return erf_approx<ERFV_EXACT>(T, x);
    case VRT_CUDA_FN_EXP: return exp_approx<EXPV_EXACT>(T, x);
    case VRT_CUDA_FN_FAST_EXP: return exp_approx<EXPV_FAST>(T, x);
    default: return exp_approx<EXPV_SPLINE>(T, x);
    }
}

template <int FN>
__global__ void __launch_bounds__(256) k_approx_rate(float *out, int iters)
{
    __shared__ ApproxTables s_tab;
    load_tables(s_tab);
    constexpr bool is_exp = FN >= VRT_CUDA_FN_EXP;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        const float u = (float)((threadIdx.x * 8 + i) % 997) / 997.f; // [0, 1)
        x[i] = is_exp ? -8.f * u : 5.f * u - 2.5f;
    }
    float acc = 0.f;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            const float y = approx_fn<FN>(s_tab, x[i]);
            acc += y;
            x[i] = fmaf(y, is_exp ? -1e-4f : 1e-4f, x[i]);
        }
    }
    if (acc == 12345.678f) out[0] = acc; // keeps the chains alive without a store in practice
}

