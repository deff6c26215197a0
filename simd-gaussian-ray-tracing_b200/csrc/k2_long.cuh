// k2_long.cuh -- K2b for lists beyond k2_band's per-warp cache: the same banded evaluation with one CTA per work item.
// A fragment of vrt_cuda.cu: included there, in this order, inside its anonymous namespace.  Not a stand-alone header.
#pragma once

// ------------------------------------------------------------------------------------------------
// K2b-long: banded render of the long lists (WIN_CAP < n <= LONG_CAP) whose band is narrow (K1 marks the cells whose band is
// more than 40 % of the list -- the bundled OBJ scenes, sigma = 0.05 in an object two units deep -- and queues them for
// k2_render<WIN>; k1_leaf in k1_bin.cuh, queue_key in k1_tile.cuh; measurements: profiles/r02_long_lists_ab.md).
// Replaces the emitter loop of src/vrt/rt.h:209-221 and the occluder loop of rt.h:107-124, like k2_band.
// ------------------------------------------------------------------------------------------------
// Everything k2_band caches per list is WARP-UNIFORM (records, the four depth bounds per entry, their running extrema), so
// it does not depend on which emitters of the list a warp works on.  Here the four warps of a CTA share ONE cache in dynamic
// shared memory (80 B per entry, sized to the frame's longest list) and work on the same 32 pixels of the same item:
//   stage    all 128 threads gather the records and compute the corner-ray bounds, one entry per thread and step
//   pass A   each warp walks a QUARTER of the list; the per-lane sums (C, sum A, emission weights) are added in warp order
//            (a pure function of the list: bands of a frame stay bit-identical to the whole frame)
//   scans    warp 0: running max of front; warp 1: suffix minima of back and of the shallowest sample depth
//   skip     T at the item's shallowest sample bounds T at every sample of the item (T is non-increasing in s): when the
//            bound of k2_band's early exit already holds there the whole item is dropped -- on an opaque object the items
//            behind the surface end here, before any emitter is touched
//   pass B   the item's emitter blocks are dealt to the warps round-robin (warp w: blocks w, w + 4, ...); each warp places its window once (the
//            prefix sum at the window's front = whole pass-A quarters + a walk inside one quarter) and then runs k2_band's
//            block loop unchanged: windows per pair group, uniform tests, moving prefix sums, early exit
//   combine  the four partial radiances are added in warp order and stored (or written to the cell's partial slot)
constexpr int LONG_CAP = 832;   // longest list cached per CTA: 65 KB of dynamic shared memory, 3 CTAs per SM (4 up to ~660 entries: the
                                // cache is sized to the frame's longest list, dispatch_k2 in vrt_cuda.cu)
constexpr int LONG_WARPS = 4;
constexpr int LONG_ENTRY_BYTES = 80;

struct LongView
{
    float4 *a, *b, *c, *fb; // as BandSmem
    float *smin1, *fmx, *bmn, *srem;
};

template <int ERF, int MINB>
__global__ void __launch_bounds__(LONG_WARPS * 32, MINB) k2_band_long(const RenderArgs args, uint32_t queue_begin, uint32_t queue_end, uint32_t cap)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ float4 s_part[LONG_WARPS][32]; // pass A per warp and lane: (C, sum A, emission weight, emission weight in front of the item)
    __shared__ float4 s_L[LONG_WARPS][32];    // partial radiance per warp and lane; .x of row 0..3 doubles as the skip test's ln T
    __shared__ uint32_t s_alive[LONG_WARPS];
    __shared__ uint32_t s_item;
    constexpr int Q = BAND_Q;
    const FrameGeom &G = args.geom;
    const int tid = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    LongView sm;
    sm.a = reinterpret_cast<float4 *>(s_dyn);
    sm.b = sm.a + cap;
    sm.c = sm.b + cap;
    sm.fb = sm.c + cap;
    sm.smin1 = reinterpret_cast<float *>(sm.fb + cap);
    sm.fmx = sm.smin1 + cap;
    sm.bmn = sm.fmx + cap;
    sm.srem = sm.bmn + cap;
    const int lx = lane & (CELL_W - 1), ly = lane >> 3;
    const float tsat = ERF == 0 ? 5.5f : EX_XMAX;
    const float esat = erf_variant<ERF>(tsat);
    const bool may_exit = args.terminate && args.scene_info[1] == 0u;
    const float exit_scale = 1.7536f * __uint_as_float(args.scene_info[0]) * (1.f / (SQRT_PI_2 * LOG2E));

    for (;;)
    {
        if (tid == 0)
        {
            uint32_t q = 0xFFFFFFFFu;
            if (!(args.abort_flag && *args.abort_flag)) q = atomicAdd(args.counter + 3, 1u) + queue_begin; // interrupted render: rt.h:244-246
            s_item = q;
        }
        __syncthreads();
        const uint32_t qi = s_item;
        if (qi >= queue_end) break;
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
        const uint32_t n = min(args.list_cnt[lid], cap); // (longer lists never reach this kernel)
        const uint32_t slot = args.cell_slot ? args.cell_slot[cell] : NO_SLOT;
        const uint32_t q_begin = slot != NO_SLOT ? slice * frame_slice(G) : 0u, q_end = slot != NO_SLOT ? min(n, q_begin + frame_slice(G)) : n;

        // ---- stage: records + the warp-uniform depth bounds from the cell's four corner rays (see k2_band) ----
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
            theta = (theta == theta) ? theta : 3.0e38f;
        }
        for (uint32_t j = tid; j < n; j += LONG_WARPS * 32)
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
            const float half = tsat * 1.0000005f / b.x + 1e-6f * big;
            const float margin = 2e-6f * big;
            sm.fb[j] = make_float4(mumax + half, mumin - half, mumax + margin, mumin - margin);
            sm.smin1[j] = mumin - 4.f * b.w;
        }
        __syncthreads();

        // shallowest sample depth of the item's emitters (uniform): where the skip test looks
        float S_item = 3.0e38f;
        for (uint32_t j = q_begin + lane; j < q_end; j += 32)
        {
            const float v = sm.smin1[j];
            S_item = fminf(S_item, (v == v) ? v : -3.0e38f); // a NaN depth never licenses a skip
        }
        S_item = warp_min_f(S_item);

        // ---- pass A: this warp's quarter of the list ----
        const uint32_t chunk = (n + LONG_WARPS - 1) / LONG_WARPS;
        {
            const uint32_t jb = min(n, (uint32_t)warp * chunk), je = min(n, jb + chunk);
            float C = 0.f, total = 0.f, etot = 0.f, pe = 0.f, ltS = 0.f;
            uint32_t n_alive = 0;
            for (uint32_t j = jb; j < je; ++j)
            {
                const float4 a = sm.a[j], b = sm.b[j], fbj = sm.fb[j];
                float mu, e;
                occluder_setup(a, b, ray, mu, e);
                const float w = b.z * e;
                etot += w;
                if (j < q_begin) pe += w;
                const bool alive = __any_sync(0xffffffffu, e > args.skip_thresh);
                const float A = alive ? w : 0.f;
                total += A;
                n_alive += alive ? 1u : 0u;
                if (fbj.w * b.x >= tsat) C = fmaf(-A, esat, C);
                else C = fmaf(A, erf_variant<ERF>(-mu * b.x), C);
                // sum_j A_j erf((S_item - mu_j) r_j): the same three cases as the window walk
                if (fbj.x <= S_item) ltS = fmaf(A, esat, ltS);
                else if (fbj.y >= S_item) ltS = fmaf(-A, esat, ltS);
                else ltS = fmaf(A, erf_variant<ERF>((S_item - mu) * b.x), ltS);
                if (!alive && lane == 0)
                {
                    sm.fb[j].x = -3.0e38f;
                    sm.fb[j].y = 3.0e38f;
                }
            }
            s_part[warp][lane] = make_float4(C, total, etot, pe);
            s_L[warp][lane].x = ltS;
            if (lane == 0) s_alive[warp] = n_alive;
        }
        __syncthreads();
        float C = 0.f, total = 0.f, etot = 0.f, pe = 0.f, ltS = 0.f;
        uint32_t n_alive = 0;
#pragma unroll
        for (int w = 0; w < LONG_WARPS; ++w)
        {
            const float4 p = s_part[w][lane];
            C += p.x;
            total += p.y;
            etot += p.z;
            pe += p.w;
            ltS += s_L[w][lane].x;
            n_alive += s_alive[w];
        }
        // ---- skip: nothing of this item can add more than TERMINATE_EPS to any channel (every warp takes the same decision) ----
        const bool skip_item = may_exit && __all_sync(0xffffffffu, ex2_approx(C - ltS) * ((etot - pe) * exit_scale * 1.01f) <= TERMINATE_EPS);
        if (!skip_item)
        {
            // running max of front (warp 0), suffix minima of back and of the shallowest sample depth (warp 1), 32 entries per step
            if (warp == 0)
            {
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
            }
            else if (warp == 1)
            {
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
            }
        }
        __syncthreads();

        float Lr = 0.f, Lg = 0.f, Lb = 0.f, La = 0.f;
        uint32_t exec = 0, sat = 0, term = 0;
        // this warp's emitter blocks: every fourth one (block w, w + 4, ...).  Neighbouring blocks cost about the same, so the four
        // warps finish together (contiguous runs left three warps waiting at the item's last barrier for 13 % of their cycles);
        // consecutive windows of a warp still overlap -- a long list's window is several times wider than 16 emitters -- so they
        // move the same way k2_band's do, only in larger steps
        constexpr uint32_t STEP = LONG_WARPS * Q;
        const uint32_t qb = q_begin + (uint32_t)warp * Q, qe = q_end;
        if (skip_item)
        {
            if (warp == 0) term = (q_end - q_begin) * n_alive;
        }
        else if (qb < qe)
        {
            auto weight_of = [&](uint32_t j) -> float {
                const float4 a = sm.a[j], b = sm.b[j];
                float mu, e;
                occluder_setup(a, b, ray, mu, e);
                return sm.fb[j].x > -1.0e38f ? b.z * e : 0.f;
            };
            auto group_range = [&](uint32_t je, bool two, float &lo, float &hi) {
                lo = sm.smin1[je];
                hi = sm.fb[je].z;
                if (two)
                {
                    lo = fminf(lo, sm.smin1[je + 1]);
                    hi = fmaxf(hi, sm.fb[je + 1].z);
                }
            };
            // ---- place the window of the first block: the saturated head [0, f) is found by a uniform scan, its per-lane weight
            // is the sum of the pass-A quarters that lie inside it plus a walk over the rest (at most a quarter of the list) ----
            uint32_t f = 0, bk = 0;
            float Pf = 0.f, Pb = 0.f;
            {
                float lo0, hi0;
                group_range(qb, qb + 1 < qe, lo0, hi0);
                if (qb + 2 < qe)
                {
                    float lo, hi;
                    group_range(qb + 2, qb + 3 < qe, lo, hi);
                    lo0 = fminf(lo0, lo);
                }
                for (;; f += 32)
                {
                    const uint32_t j = f + lane;
                    const uint32_t m = __ballot_sync(0xffffffffu, j >= n || !(sm.fmx[j] <= lo0));
                    if (m) { f += __ffs(m) - 1; break; }
                }
                const uint32_t v = f / chunk;
                for (uint32_t u = 0; u < v; ++u) Pf += s_part[u][lane].y;
                for (uint32_t j = v * chunk; j < f; ++j) Pf += weight_of(j);
                bk = f;
                Pb = Pf;
            }
            uint32_t exit_check_at = qb;
            for (uint32_t q0 = qb; q0 < qe; q0 += STEP)
            {
                const uint32_t n_real = min((uint32_t)Q, qe - q0);
                float s[Q][5], acc[Q][5], wgt[Q];
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
                if (!__any_sync(0xffffffffu, any_emit)) continue;
                const bool g1 = n_real > 2;
                const uint32_t ng0 = min(2u, n_real), ng1 = n_real - ng0;
                float smin_g[2], smax_g[2];
                group_range(q0, n_real > 1, smin_g[0], smax_g[0]);
                smin_g[1] = smin_g[0];
                smax_g[1] = smax_g[0];
                if (g1) group_range(q0 + 2, n_real > 3, smin_g[1], smax_g[1]);
                const float Smin = fminf(smin_g[0], smin_g[1]), Smax = fmaxf(smax_g[0], smax_g[1]);
                float Smin_next = 3.0e38f;
                if (q0 + STEP < qe)
                {
                    float lo, hi;
                    group_range(q0 + STEP, q0 + STEP + 1 < qe, Smin_next, hi);
                    if (q0 + STEP + 2 < qe)
                    {
                        group_range(q0 + STEP + 2, q0 + STEP + 3 < qe, lo, hi);
                        Smin_next = fminf(Smin_next, lo);
                    }
                }
                // explicit window steps (rare: the window moves with the blocks, see k2_band)
                while (f < n && sm.fmx[f] <= Smin) { Pf += weight_of(f); ++f; }
                while (f > 0 && !(sm.fmx[f - 1] <= Smin)) { --f; Pf -= weight_of(f); }
                while (bk > 0 && sm.bmn[bk - 1] >= Smax) { --bk; Pb -= weight_of(bk); }
                while (bk < f) { Pb += weight_of(bk); ++bk; }
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
                    if (!(fbj.x > -1.0e38f)) continue;
                    const float4 a = sm.a[j], b = sm.b[j];
                    float mu, e;
                    occluder_setup(a, b, ray, mu, e);
                    const float A = b.z * e, r = b.x, nm = -(mu - s0) * r;
                    if (j >= bk) Pb += A;
                    if (j < f_next) Pf += A;
                    const float2 rr = make_float2(r, r), mm = make_float2(nm, nm);
#pragma unroll
                    for (int g = 0; g < 2; ++g)
                    {
                        if (g == 1 && !g1) break;
                        const float sming = smin_g[g], smaxg = smax_g[g];
                        float &base = g == 0 ? base0 : base1;
                        const uint32_t ng = g == 0 ? ng0 : ng1;
                        if (fbj.x <= sming) { base = fmaf(A, esat, base); continue; }
                        if (fbj.y >= smaxg) { base = fmaf(-A, esat, base); continue; }
                        exec += ng;
                        const bool pos = fbj.z <= sming;
                        const bool neg = fbj.w >= smaxg;
                        if (pos || neg)
                        {
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
                const float base_common = esat * (Pf_now - (total - Pb));
                base0 += base_common;
                base1 += base_common;
                sat += n_alive * n_real;
                float lt_max = -3.0e38f;
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
                // early termination of this warp's remaining blocks (k2_band's test; `pe` lacks the other warps' emitters in front:
                // conservative; S is the shallowest sample of ALL emitters behind this block, this warp's among them)
                if (may_exit && q0 + STEP < qe && q0 >= exit_check_at)
                {
                    const float w_rem = (etot - pe) * exit_scale * 1.01f;
                    if (__all_sync(0xffffffffu, ex2_approx(lt_max) * w_rem <= TERMINATE_EPS))
                    {
                        const float S = sm.srem[q0 + Q];
                        float lt = C;
                        for (uint32_t j = 0; j < n; ++j)
                        {
                            const float4 fbj = sm.fb[j];
                            if (!(fbj.x > -1.0e38f)) continue;
                            const float4 a = sm.a[j], b = sm.b[j];
                            float mu, e;
                            occluder_setup(a, b, ray, mu, e);
                            const float A = b.z * e;
                            if (fbj.x <= S) lt = fmaf(-A, esat, lt);
                            else if (fbj.y >= S) lt = fmaf(A, esat, lt);
                            else lt = fmaf(-A, erf_variant<ERF>((S - mu) * b.x), lt);
                        }
                        if (__all_sync(0xffffffffu, ex2_approx(lt) * w_rem <= TERMINATE_EPS))
                        {
                            for (uint32_t qr = q0 + STEP; qr < qe; qr += STEP) term += min((uint32_t)Q, qe - qr) * n_alive; // this warp's remaining emitters
                            break;
                        }
                        exit_check_at = q0 + 3 * STEP;
                    }
                }
            }
        }
        // ---- combine the four runs in warp order ----
        s_L[warp][lane] = make_float4(Lr, Lg, Lb, La);
        if (lane == 0 && exec) atomicAdd(args.terms_exec, (unsigned long long)exec * 5ull * n_live);
        if (lane == 0 && sat > exec) atomicAdd(args.terms_sat, (unsigned long long)(sat - exec) * 5ull * n_live);
        if (lane == 0 && term) atomicAdd(args.terms_term, (unsigned long long)term * 5ull * n_live);
        __syncthreads();
        if (warp == 0)
        {
#pragma unroll
            for (int w = 1; w < LONG_WARPS; ++w)
            {
                const float4 p = s_L[w][lane];
                Lr += p.x;
                Lg += p.y;
                Lb += p.z;
                La += p.w;
            }
            if (slot != NO_SLOT) args.partial[(size_t)(slot + slice) * 32 + lane] = make_float4(Lr, Lg, Lb, La); // summed by k3_combine
            else store_cell(args, G, px, py, live, Lr, Lg, Lb, La);
        }
    }
}
