// k2_render.cuh -- K2 + K3: the render kernel (one warp per 8x4-pixel cell), pixel packing, the combine pass of split cells.
// A fragment of vrt_cuda.cu: included there, in this order, inside its anonymous namespace.  Not a stand-alone header.
#pragma once

// ------------------------------------------------------------------------------------------------
// K2 + K3: render
// ------------------------------------------------------------------------------------------------
struct RenderArgs
{
    FrameGeom geom;            // the frame (by value: nothing about a frame lives in per-device state)
    const Rec *rec;            // frame records
    const uint32_t *list_off;  // per list: entries [off, off + cnt) of list_idx (or of rec when the lists are contiguous)
    const uint32_t *list_cnt;
    const uint32_t *list_idx;  // indices into rec, or nullptr when lists are contiguous ranges of rec
    const uint32_t *queue;     // cost-ordered cell ids
    uint32_t n_queue;
    uint32_t *counter;         // work-queue head
    uint32_t *image;           // W*H packed pixels (may be null)
    float4 *radiance;          // W*H float4 (may be null)
    unsigned long long *terms_exec;
    unsigned long long *terms_sat;
    unsigned long long *terms_term;  // terms dropped by the transmittance early exit (k2_band)
    const uint32_t *scene_info;      // K0: [0] max |albedo| (float bits), [1] != 0: non-monotone scene, no early exit
    const volatile uint32_t *abort_flag; // mapped host word, or null: != 0 -> the persistent warps stop taking work (`running` went false)
    uint32_t terminate;              // early exit enabled
    float skip_thresh;         // skip an occluder / emitter for the whole warp when exp2 weight <= thresh (-1: never)
    uint32_t quant_nearest, alpha_from_w;
    uint32_t image_vec16;      // the image base and row pitch are 16-byte aligned: quads of pixels are stored as one uint4
    uint32_t window; // banded evaluation of depth-sorted lists (k2_band; k2_render<WIN> for lists beyond its cache)
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

// K3: clamp, quantise (truncate | round-to-nearest-even) and pack one pixel (rt.h:238-243 / 329-333, alpha quirk rt.h:373-377)
__device__ __forceinline__ uint32_t pack_pixel(const RenderArgs &args, float Lr, float Lg, float Lb, float La)
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
    return (A << 24) | (R << 16) | (Gc << 8) | B;
}

// Framebuffer write of one 8x4-pixel cell by its warp (lane = pixel, lane = ly * 8 + lx); EVERY lane calls it.
// The packed words of a row are transposed through three shuffles so that two lanes per row store 16 bytes each: eight
// 128-bit stores per cell instead of 32 scalar ones (replaces the per-pixel stores and the tile -> image copy of
// rt.h:373-377, 388-399).  Rows that are not entirely inside the image / band, or a frame whose rows are not 16-byte
// aligned, fall back to one word per lane.  The optional float4 radiance is one 16-byte store per lane already.
__device__ __forceinline__ void store_cell(const RenderArgs &args, const FrameGeom &G, int px, int py, bool live, float Lr, float Lg, float Lb, float La)
{
    const size_t pi = (size_t)py * G.W + px;
    if (args.radiance && live) args.radiance[pi] = make_float4(Lr, Lg, Lb, La);
    if (!args.image) return;
    const int lane = threadIdx.x & 31;
    const uint32_t w0 = pack_pixel(args, Lr, Lg, Lb, La);
    const uint32_t w1 = __shfl_down_sync(0xffffffffu, w0, 1), w2 = __shfl_down_sync(0xffffffffu, w0, 2), w3 = __shfl_down_sync(0xffffffffu, w0, 3);
    // the four lanes of a quad are live together, their pixels are consecutive and the first one is 16-byte aligned
    const uint32_t live_mask = __ballot_sync(0xffffffffu, live);
    const bool quad = ((live_mask >> (lane & ~3)) & 0xFu) == 0xFu && (pi & 3) == 0 && args.image_vec16;
    const bool quad_any = __shfl_sync(0xffffffffu, (int)quad, lane & ~3) != 0;
    if (quad_any)
    {
        if ((lane & 3) == 0) *reinterpret_cast<uint4 *>(args.image + pi) = make_uint4(w0, w1, w2, w3);
    }
    else if (live) args.image[pi] = w0;
}

// ray through pixel (px, py): plane = inverse(view) (u, v, 0, 1) (src/vrt/camera.cpp:60-70); dir = normalize(plane - origin)
__device__ __forceinline__ PixelRay pixel_ray(const FrameGeom &G, int px, int py)
{
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
    const FrameGeom &G = args.geom;
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
        if (args.abort_flag && *args.abort_flag) break; // interrupted render (`running` went false, rt.h:244-246, 289)
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
        const uint32_t q_begin = slot != NO_SLOT ? slice * frame_slice(G) : 0u, q_end = slot != NO_SLOT ? min(n, q_begin + frame_slice(G)) : n;
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
        else if (live)
        {
            // One word per lane (32-byte row segments) in THIS kernel: with the 16-byte form inlined into its epilogue ptxas
            // schedules the inner term 6 % slower (149.2 vs 140.9 ms on BASELINE config 5, same box, profiles/r02_strict_ab.md);
            // the default kernel (k2_band), the combine pass and the variants store 16 bytes.
            const size_t pi = (size_t)py * G.W + px;
            if (args.radiance) args.radiance[pi] = make_float4(Lr, Lg, Lb, La);
            if (args.image) args.image[pi] = pack_pixel(args, Lr, Lg, Lb, La);
        }
        if (lane == 0 && exec) atomicAdd(args.terms_exec, exec * 5ull * n_live);
        if (WIN && lane == 0 && sat) atomicAdd(args.terms_sat, sat * 5ull * n_live);
        exec = 0;
        sat = 0;
    }
}

// K3', split cells: sum the slices' partial radiances in slice order (deterministic) and write the pixel.
__global__ void __launch_bounds__(256) k3_combine(const RenderArgs args, int cy_begin, int cy_end)
{
    const FrameGeom &G = args.geom;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t ncells = (uint32_t)((cy_end - cy_begin) * G.ncx);
    if (w >= ncells) return;
    const int cx = (int)(w % G.ncx), cy = cy_begin + (int)(w / G.ncx);
    const uint32_t cell = (uint32_t)(cy * G.ncx + cx);
    int x0, y0, cw, ch;
    cell_rect(G, cx, cy, x0, y0, cw, ch);
    if (min(y0 + ch, G.row_end) <= max(y0, G.row_begin)) return; // not queued: its slot entry is stale
    const uint32_t slot = args.cell_slot[cell];
    if (slot == NO_SLOT) return;
    const uint32_t lid = cell_list_id(G, cx, cy);
    const uint32_t items = cell_items(G, args.list_cnt[lid], cell);
    float4 L = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t k = 0; k < items; ++k)
    {
        const float4 p = args.partial[(size_t)(slot + k) * 32 + lane];
        L.x += p.x; L.y += p.y; L.z += p.z; L.w += p.w;
    }
    const int lx = lane & (CELL_W - 1), ly = lane >> 3;
    const int px = x0 + lx, py = y0 + ly;
    store_cell(args, G, px, py, lx < cw && ly < ch && py >= G.row_begin && py < G.row_end, L.x, L.y, L.z, L.w);
}
