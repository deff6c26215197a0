// k2_band.cuh -- K2b: the default render kernel for depth-sorted per-cell lists (banded evaluation + early termination).
// A fragment of vrt_cuda.cu: included there, in this order, inside its anonymous namespace.  Not a stand-alone header.
#pragma once

// ------------------------------------------------------------------------------------------------
// K2b: banded render.  Replaces the emitter loop of src/vrt/rt.h:209-221 and the occluder loop of rt.h:107-124.
// ------------------------------------------------------------------------------------------------
// Lists are depth-sorted by K1 (k1_sort_cells).  In depth order the (emitter q, occluder j) interaction matrix is BANDED:
// erf((s - mu_j) r_j) is saturated to +-esat in fp32 unless the sample depth s lies within t_sat standard widths of
// mu_j, so only occluders near the emitter's own depth need the erf evaluated; everything in front contributes +A_j esat,
// everything behind -A_j esat.  Per cell (one warp, lane = pixel):
//   pass A   one walk over the list: C = sum_j A_j erf(-m_j), total = sum_j A_j, and per occluder four WARP-UNIFORM depths
//            (front_j, back_j, mumax_j, mumin_j): beyond front_j every lane's argument is >= t_sat, before back_j it is
//            <= -t_sat; plus their running max / suffix min (so the saturated head and tail of the list are intervals) and
//            the suffix min of the shallowest sample depth of the remaining emitters (for the early exit).
//   pass B   emitters in register blocks of 4 = two PAIR GROUPS (a packed f32x2 lane pair = two emitters x 5 samples).
//            Per block the head [0, f) and tail [bk, n) of the list are resolved by two running per-lane prefix sums that
//            move with the block; the window [f, bk) is walked once, and every window occluder is tested per pair group
//            with uniform compares only (no per-lane votes): saturated in front / behind -> one FFMA into the group's base,
//            sign-uniform -> 8 packed FMA-pipe ops + 1 MUFU per term, otherwise the signed body.  The window of a PAIR is
//            less than half the window of an 8-emitter block, which is where the speed-up over the round-1 depth-window
//            kernel comes from (2.3x fewer evaluated terms on BASELINE config 5).
//   exit     T(s) is non-increasing in s (every A_j >= 0, erf increasing).  After a block whose own samples are all dark for
//            every lane (a cheap predictor), T is evaluated once at S_rem, the shallowest sample depth of ALL remaining
//            emitters; if T(S_rem) * (remaining emission weight) <= eps for every lane, the remaining emitters change no
//            channel by more than eps and the warp breaks out of the emitter loop.  Disabled when the scene
//            holds a negative magnitude (K0 flags it) or with VRT_CUDA_NO_TERMINATE.
constexpr int BAND_Q = 4;
constexpr int BAND_WARPS = 4;
constexpr float TERMINATE_EPS = 1e-6f; // bound on the per-channel radiance dropped by the early exit

struct BandSmem
{
    float4 a[WIN_CAP], b[WIN_CAP]; // occluder part of the records
    float4 c[WIN_CAP];             // albedo (read at a block's epilogue: keeps 16 registers free during the window walk)
    float4 fb[WIN_CAP];            // (front_j, back_j, mumax_j + margin, mumin_j - margin); front = -3e38: no lane sees it

    float smin1[WIN_CAP];          // mumin_j - 4 sigma_j: shallowest sample depth of emitter j over the warp
    float fmx[WIN_CAP];            // running max of front  : j < f  <=>  fmx[j] <= Smin
    float bmn[WIN_CAP];            // suffix  min of back   : j >= bk <=>  bmn[j] >= Smax
    float srem[WIN_CAP];           // suffix  min of mumin - 4 sigma: shallowest sample of the emitters j, j+1, ...
};

__device__ __forceinline__ int ordered_int(float x)
{
    const int k = __float_as_int(x);
    return k ^ ((k >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ordered_float(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
__device__ __forceinline__ float warp_max_f(float x) { return ordered_float(__reduce_max_sync(0xffffffffu, ordered_int(x))); }
__device__ __forceinline__ float warp_min_f(float x) { return ordered_float(__reduce_min_sync(0xffffffffu, ordered_int(x))); }

template <int ERF, int MINB>
__global__ void __launch_bounds__(BAND_WARPS * 32, MINB) k2_band(const RenderArgs args, uint32_t queue_begin)
{
    __shared__ BandSmem s_band[BAND_WARPS];
    constexpr int Q = BAND_Q;
    const FrameGeom &G = args.geom;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    BandSmem &sm = s_band[warp];
    const int lx = lane & (CELL_W - 1), ly = lane >> 3;
    const float tsat = ERF == 0 ? 5.5f : EX_XMAX;
    const float esat = erf_variant<ERF>(tsat);
    // early exit: radiance dropped <= T_bound * sum(remaining sigma c_bar) * sum_k e^{-k^2/2} * max |albedo|
    const bool may_exit = args.terminate && args.scene_info[1] == 0u;
    const float exit_scale = 1.7536f * __uint_as_float(args.scene_info[0]) * (1.f / (SQRT_PI_2 * LOG2E));

    for (;;)
    {
        if (args.abort_flag && *args.abort_flag) break; // interrupted render (`running` went false, rt.h:244-246)
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(args.counter + 1, 1u) + queue_begin;
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
        const uint32_t n = min(args.list_cnt[lid], (uint32_t)WIN_CAP); // (longer lists never reach this kernel)
        const uint32_t slot = args.cell_slot ? args.cell_slot[cell] : NO_SLOT;
        const uint32_t q_begin = slot != NO_SLOT ? slice * frame_slice(G) : 0u, q_end = slot != NO_SLOT ? min(n, q_begin + frame_slice(G)) : n;

        // The depth bounds of pass A are WARP-UNIFORM, and every ray of the cell lies inside the pyramid of its four corner
        // rays: min over the lanes of mu = oc . n is bounded below by the minimum over the corner rays (the direction farthest
        // from oc within a convex set of directions is a vertex), the maximum above by the corner maximum plus
        // (k sigma + |oc| theta) theta + |oc| theta^2 / 2 (theta = angular diagonal of the cell; the centre of a listed Gaussian
        // lies within k sigma of the cell's frustum, so its distance to any of the cell's rays is at most k sigma + |oc| theta).
        // So they are computed by the lanes in parallel over the ENTRIES while staging, not by warp reductions per entry.
        float crx[4], cry[4], crz[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
        {
            const int src = (c & 1 ? 7 : 0) + (c & 2 ? 24 : 0);
            crx[c] = __shfl_sync(0xffffffffu, ray.nx, src);
            cry[c] = __shfl_sync(0xffffffffu, ray.ny, src);
            crz[c] = __shfl_sync(0xffffffffu, ray.nz, src);
        }
        float theta;
        {
            const float dx = crx[0] - crx[3], dy = cry[0] - cry[3], dz = crz[0] - crz[3];
            theta = sqrtf(dx * dx + dy * dy + dz * dz) * 1.0001f + 1e-7f;
            theta = (theta == theta) ? theta : 3.0e38f; // a degenerate ray in the cell: no saturation shortcut anywhere
        }
        __syncwarp();
        for (uint32_t j = lane; j < n; j += 32)
        {
            const Rec *r = args.rec + args.list_idx[off + j];
            const float4 a = r->a, b = r->b;
            sm.a[j] = a;
            sm.b[j] = b;
            sm.c[j] = r->c;
            float mumin = 3.0e38f, mumax = -3.0e38f;
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
                const float m = fmaf(a.z, crz[c], fmaf(a.y, cry[c], a.x * crx[c]));
                mumin = fminf(mumin, m);
                mumax = fmaxf(mumax, m);
            }
            const float ocn = sqrtf(fmaf(a.z, a.z, fmaf(a.y, a.y, a.x * a.x)));
            mumax += (G.bound_k * b.w + 1.5f * ocn * theta) * theta * 1.0001f;
            const float big = fmaxf(fabsf(mumax), fabsf(mumin));
            const float half = tsat * 1.0000005f / b.x + 1e-6f * big; // t >= tsat must hold after fp32 rounding of t
            const float margin = 2e-6f * big;                        // sign tests: a sample within a few ulp of the centre takes the signed body
            sm.fb[j] = make_float4(mumax + half, mumin - half, mumax + margin, mumin - margin);
            sm.smin1[j] = mumin - 4.f * b.w;
        }
        __syncwarp();

        // ---- pass A ----
        float C = 0.f, total = 0.f, etot = 0.f, pe = 0.f;
        uint32_t n_alive = 0;
        {
            for (uint32_t j = 0; j < n; ++j)
            {
                const float4 a = sm.a[j], b = sm.b[j];
                float mu, e;
                occluder_setup(a, b, ray, mu, e);
                const float w = b.z * e;
                etot += w;                 // emission weights keep every entry
                if (j < q_begin) pe += w;  // ... of the emitters in front of this item's range
                // an occluder no lane sees (weight <= threshold everywhere) is dropped exactly like the plain kernel's skip
                const bool alive = __any_sync(0xffffffffu, e > args.skip_thresh);
                const float A = alive ? w : 0.f;
                total += A;
                n_alive += alive ? 1u : 0u;
                // erf(-m) is saturated for an occluder more than t_sat widths beyond the origin (the usual case): same value, no erf
                if (sm.fb[j].w * b.x >= tsat) C = fmaf(-A, esat, C);
                else C = fmaf(A, erf_variant<ERF>(-mu * b.x), C);
                if (!alive && lane == 0)
                {
                    sm.fb[j].x = -3.0e38f; // never blocks the saturated head ...
                    sm.fb[j].y = 3.0e38f;  // ... nor the tail
                }
            }
            __syncwarp();
            // running max of front (forwards), suffix minima of back and of the shallowest sample depth (backwards), 32 entries per step
            float carry_f = -3.0e38f;
            for (uint32_t j0 = 0; j0 < n; j0 += 32)
            {
                const uint32_t j = j0 + lane;
                float vf = j < n ? sm.fb[j].x : -3.0e38f;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1)
                {
                    const float t = __shfl_up_sync(0xffffffffu, vf, d);
                    if (lane >= d) vf = fmaxf(vf, t);
                }
                vf = fmaxf(vf, carry_f);
                if (j < n) sm.fmx[j] = vf;
                carry_f = __shfl_sync(0xffffffffu, vf, 31);
            }
            float carry_b = 3.0e38f, carry_s = 3.0e38f;
            for (int j0 = (int)((n - 1) & ~31u); n && j0 >= 0; j0 -= 32)
            {
                const uint32_t j = (uint32_t)j0 + lane;
                float vb = 3.0e38f, vs = 3.0e38f;
                if (j < n)
                {
                    vb = sm.fb[j].y;
                    vs = sm.smin1[j];
                    vs = (vs == vs) ? vs : -3.0e38f; // a NaN depth never licenses an exit
                }
#pragma unroll
                for (int d = 1; d < 32; d <<= 1)
                {
                    const float tb = __shfl_down_sync(0xffffffffu, vb, d), ts = __shfl_down_sync(0xffffffffu, vs, d);
                    if (lane + d < 32)
                    {
                        vb = fminf(vb, tb);
                        vs = fminf(vs, ts);
                    }
                }
                vb = fminf(vb, carry_b);
                vs = fminf(vs, carry_s);
                if (j < n)
                {
                    sm.bmn[j] = vb;
                    sm.srem[j] = vs;
                }
                carry_b = __shfl_sync(0xffffffffu, vb, 0);
                carry_s = __shfl_sync(0xffffffffu, vs, 0);
            }
            __syncwarp();
        }

        // per-lane weight of occluder j for the running prefix sums (0 for an entry no lane sees)
        auto weight_of = [&](uint32_t j) -> float {
            const float4 a = sm.a[j], b = sm.b[j];
            float mu, e;
            occluder_setup(a, b, ray, mu, e);
            return sm.fb[j].x > -1.0e38f ? b.z * e : 0.f;
        };

        // ---- pass B ----
        float Lr = 0.f, Lg = 0.f, Lb = 0.f, La = 0.f;
        uint32_t exec = 0, sat = 0, term = 0; // per item: at most n x emitters (x 5 n_live at the flush)
        // window [f, bk) of the current block; Pf = sum_{j<f} A_j, Pb = sum_{j<bk} A_j (per lane); nf / nb count the entries
        // some lane sees among [0, f) / [0, bk).  The window moves with the blocks WITHOUT extra per-lane work in the common case:
        // an occluder that enters at the back is evaluated by the block it enters in (its weight is added to Pb there), and one
        // that leaves at the front was in the previous block's window, where the next block's shallowest sample depth was
        // already known (its weight was added to Pf there).  The explicit loops below only run at an item's first block, when
        // the window has to move backwards, or when consecutive windows do not overlap.
        uint32_t f = 0, bk = 0;
        float Pf = 0.f, Pb = 0.f;
        uint32_t exit_check_at = q_begin; // first block after which the exit test may run again (back-off after a failed test)
        // uniform sample-depth range of the emitter pair (je, je + 1 if real): [min(mumin - 4 sigma), max(mumax)]
        auto group_range = [&](uint32_t je, bool two, float &lo, float &hi) {
            lo = sm.smin1[je];
            hi = sm.fb[je].z;
            if (two)
            {
                lo = fminf(lo, sm.smin1[je + 1]);
                hi = fmaxf(hi, sm.fb[je + 1].z);
            }
        };
        for (uint32_t q0 = q_begin; q0 < q_end; q0 += Q)
        {
            const uint32_t n_real = min((uint32_t)Q, q_end - q0);
            float s[Q][5], acc[Q][5], wgt[Q];
            // one depth shift per block keeps s r - m small (uniform: the first emitter's shallowest centre depth)
            float s0 = sm.fb[q0].w;
            s0 = (fabsf(s0) <= 3.0e38f) ? s0 : 0.f;
            bool any_emit = false;
#pragma unroll
            for (int e = 0; e < Q; ++e)
            {
                const bool real = (uint32_t)e < n_real;
                const uint32_t je = real ? q0 + e : q0;
                const float4 a = sm.a[je], b = sm.b[je];
                float mu, ee;
                occluder_setup(a, b, ray, mu, ee);
                const float w = b.z * ee;
                wgt[e] = real ? w * (1.f / (SQRT_PI_2 * LOG2E)) : 0.f;
                if (real) pe += w;
                any_emit |= real && (ee > args.skip_thresh);
#pragma unroll
                for (int k = 0; k < 5; ++k)
                {
                    s[e][k] = (mu - s0) + (float)(k - 4) * b.w;
                    acc[e][k] = 0.f;
                }
            }
            if (!__any_sync(0xffffffffu, any_emit)) continue; // no lane sees any of these emitters
            const bool g1 = n_real > 2;
            const uint32_t ng0 = min(2u, n_real), ng1 = n_real - ng0;
            float smin_g[2], smax_g[2];
            group_range(q0, n_real > 1, smin_g[0], smax_g[0]);
            smin_g[1] = smin_g[0];
            smax_g[1] = smax_g[0];
            if (g1) group_range(q0 + 2, n_real > 3, smin_g[1], smax_g[1]);
            const float Smin = fminf(smin_g[0], smin_g[1]), Smax = fmaxf(smax_g[0], smax_g[1]);
            // the next block's shallowest sample depth: what leaves the window after this block
            float Smin_next = 3.0e38f;
            if (q0 + Q < q_end)
            {
                float lo, hi;
                group_range(q0 + Q, q0 + Q + 1 < q_end, Smin_next, hi);
                if (q0 + Q + 2 < q_end)
                {
                    group_range(q0 + Q + 2, q0 + Q + 3 < q_end, lo, hi);
                    Smin_next = fminf(Smin_next, lo);
                }
            }

            // ---- place the window (explicit steps: rare, see above) ----
            while (f < n && sm.fmx[f] <= Smin) { Pf += weight_of(f); ++f; }
            while (f > 0 && !(sm.fmx[f - 1] <= Smin)) { --f; Pf -= weight_of(f); }
            while (bk > 0 && sm.bmn[bk - 1] >= Smax) { --bk; Pb -= weight_of(bk); }
            while (bk < f) { Pb += weight_of(bk); ++bk; } // (an empty window in front of f)
            // two uniform scans, no per-lane work: the tail [bk_new, n) is behind every sample of this block, and the window
            // entries [f, f_next) will be in front of every sample of the next one
            // (32 entries per step, one per lane: a window rarely spans more)
            uint32_t bk_new = bk, f_next = f;
            for (;; bk_new += 32)
            {
                const uint32_t j = bk_new + lane;
                const uint32_t m = __ballot_sync(0xffffffffu, j >= n || sm.bmn[j] >= Smax);
                if (m) { bk_new += __ffs(m) - 1; break; }
            }
            for (;; f_next += 32)
            {
                const uint32_t j = f_next + lane;
                const uint32_t m = __ballot_sync(0xffffffffu, j >= bk_new || !(sm.fmx[j] <= Smin_next));
                if (m) { f_next += __ffs(m) - 1; break; }
            }

            const float Pf_now = Pf;
            const uint32_t f_now = f;
            float base0 = 0.f, base1 = 0.f;
            for (uint32_t j = f_now; j < bk_new; ++j)
            {
                const float4 fbj = sm.fb[j];
                if (!(fbj.x > -1.0e38f)) continue; // no lane sees it: weight 0 everywhere
                const float4 a = sm.a[j], b = sm.b[j];
                float mu, e;
                occluder_setup(a, b, ray, mu, e);
                const float A = b.z * e, r = b.x, nm = -(mu - s0) * r;
                if (j >= bk) Pb += A;     // enters the window at the back
                if (j < f_next) Pf += A;  // leaves it at the front after this block
                const float2 rr = make_float2(r, r), mm = make_float2(nm, nm);
#pragma unroll
                for (int g = 0; g < 2; ++g)
                {
                    if (g == 1 && !g1) break;
                    const float sming = smin_g[g], smaxg = smax_g[g];
                    float &base = g == 0 ? base0 : base1;
                    const uint32_t ng = g == 0 ? ng0 : ng1;
                    if (fbj.x <= sming) { base = fmaf(A, esat, base); continue; }  // in front of every sample: erf = +esat
                    if (fbj.y >= smaxg) { base = fmaf(-A, esat, base); continue; } // behind every sample: -esat
                    exec += ng;
                    const bool pos = fbj.z <= sming; // every sample behind the centre for every lane: t >= 0
                    const bool neg = fbj.w >= smaxg; // every sample in front of it: t <= 0
                    if (pos || neg)
                    {
                        // erf(t) = +-(1 - w(t)), sign known per (occluder, group): +-A once, -+A w(t) per term
                        base += pos ? A : -A;
                        const float sA = pos ? -A : A;
                        const float2 AA = make_float2(sA, sA);
#pragma unroll
                        for (int k = 0; k < 5; ++k)
                        {
                            const float2 t = __ffma2_rn(make_float2(s[2 * g][k], s[2 * g + 1][k]), rr, mm);
                            const float2 ac = __ffma2_rn(AA, erfc_mag2<ERF>(t), make_float2(acc[2 * g][k], acc[2 * g + 1][k]));
                            acc[2 * g][k] = ac.x;
                            acc[2 * g + 1][k] = ac.y;
                        }
                    }
                    else
                    {
                        const float2 AA = make_float2(A, A);
#pragma unroll
                        for (int k = 0; k < 5; ++k)
                        {
                            const float2 t = __ffma2_rn(make_float2(s[2 * g][k], s[2 * g + 1][k]), rr, mm);
                            const float2 ac = __ffma2_rn(AA, erf_variant2<ERF>(t), make_float2(acc[2 * g][k], acc[2 * g + 1][k]));
                            acc[2 * g][k] = ac.x;
                            acc[2 * g + 1][k] = ac.y;
                        }
                    }
                }
            }
            bk = bk_new;
            f = f_next;
            // head [0, f_now) in front of every sample, tail [bk, n) behind: resolved by the two prefix sums
            const float base_common = esat * (Pf_now - (total - Pb));
            base0 += base_common;
            base1 += base_common;
            sat += n_alive * n_real; // every entry some lane sees is, for each emitter of the block, either evaluated or saturated (exec is subtracted at the flush)
            // T(s) = 2^(C - base - acc); pdf at the samples = c_bar e^{-k^2/2}, k = -4..0 (src/vrt/rt.h:153-161)
            float lt_max = -3.0e38f; // log2 T at the block's least occluded sample (k = -4 of every emitter)
#pragma unroll
            for (int e = 0; e < Q; ++e)
            {
                const float Cb = C - ((e >> 1) == 0 ? base0 : base1);
                const float l0 = Cb - acc[e][0];
                float inner = 3.3546262790251185e-4f * ex2_approx(l0);
                inner = fmaf(1.1108996538242306e-2f, ex2_approx(Cb - acc[e][1]), inner);
                inner = fmaf(1.3533528323661270e-1f, ex2_approx(Cb - acc[e][2]), inner);
                inner = fmaf(6.0653065971263342e-1f, ex2_approx(Cb - acc[e][3]), inner);
                inner += ex2_approx(Cb - acc[e][4]);
                inner *= wgt[e];
                const float4 al = sm.c[(uint32_t)e < n_real ? q0 + e : q0];
                Lr = fmaf(al.x, inner, Lr);
                Lg = fmaf(al.y, inner, Lg);
                Lb = fmaf(al.z, inner, Lb);
                La = fmaf(al.w, inner, La);
                if ((uint32_t)e < n_real) lt_max = fmaxf(lt_max, l0);
            }
            // ---- early termination: nothing behind can add more than TERMINATE_EPS to any channel ----
            if (may_exit && q0 + Q < q_end && q0 >= exit_check_at)
            {
                const float w_rem = (etot - pe) * exit_scale * 1.01f; // (1 % for the fp32 rounding of the bound itself)
                // predictor: this block's own samples are already dark for every lane
                if (__all_sync(0xffffffffu, ex2_approx(lt_max) * w_rem <= TERMINATE_EPS))
                {
                    // T at the shallowest sample depth of all remaining emitters bounds T at every sample still to come
                    const float S = sm.srem[q0 + Q];
                    float lt = C;
                    for (uint32_t j = 0; j < n; ++j)
                    {
                        const float4 a = sm.a[j], b = sm.b[j];
                        float mu, e;
                        occluder_setup(a, b, ray, mu, e);
                        const float A = sm.fb[j].x > -1.0e38f ? b.z * e : 0.f;
                        lt = fmaf(-A, erf_variant<ERF>((S - mu) * b.x), lt);
                    }
                    if (__all_sync(0xffffffffu, ex2_approx(lt) * w_rem <= TERMINATE_EPS))
                    {
                        term += (q_end - (q0 + Q)) * n_alive;
                        break;
                    }
                    exit_check_at = q0 + 3 * Q; // not yet: look again two blocks further on
                }
            }
        }
        if (slot != NO_SLOT) args.partial[(size_t)(slot + slice) * 32 + lane] = make_float4(Lr, Lg, Lb, La); // summed by k3_combine
        // 16-byte framebuffer stores (shuffle transpose): north star (3), and what the zero-copy output wants (wide PCIe writes).  On
        // BASELINE config 5 they cost nothing against one word per lane (20.611 vs 20.605 ms, profiles/r02_band_ab.md).
        else store_cell(args, G, px, py, live, Lr, Lg, Lb, La);
        if (lane == 0 && exec) atomicAdd(args.terms_exec, (unsigned long long)exec * 5ull * n_live);
        if (lane == 0 && sat > exec) atomicAdd(args.terms_sat, (unsigned long long)(sat - exec) * 5ull * n_live);
        if (lane == 0 && term) atomicAdd(args.terms_term, (unsigned long long)term * 5ull * n_live);
    }
}
