// k2_window.cuh -- K2': the depth-window render kernel.
// A fragment of vrt_cuda.cu: included there, in this order, inside its anonymous namespace (one translation unit, so
// every kernel sees the same __constant__ frame geometry).  Not a stand-alone header.
#pragma once

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
