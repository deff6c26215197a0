// k1_bin.cuh -- K1 for the bounded list modes: per-Gaussian screen binning + one fused cull / depth-sort pass per cell.
// A fragment of vrt_cuda.cu: included there, in this order, inside its anonymous namespace.  Not a stand-alone header.
#pragma once

// ------------------------------------------------------------------------------------------------
// Replaces the O(tiles x N) scan of src/vrt/rt.cpp:47-66 (and round 1's four-level hierarchy, which still scanned
// groups x segments at its root) by per-Gaussian tile-range binning:
//   k1_bin_count   one thread per Gaussian: the rectangle of BINS (4 x 8 cells = 32 x 32 pixels) whose ray frustum can come
//                  within k sigma of the centre -- two closed-form intervals, one per screen axis -- clipped to the rendered
//                  row band; one atomic per covered bin.  A rank that renders 1/8 of the frame touches 1/8 of the bins.
//   (scan of the bin counts)
//   k1_bin_fill    the same rectangles again, every Gaussian appends itself to its bins (atomic cursor per bin)
//   k1_leaf        one warp per 8x4-pixel cell: scans its bin's list ONCE with the exact predicate (reference cull test of
//                  rt.cpp:57-59 AND the rounded-corner k-sigma frustum bound, warp ballot + popc compaction into shared
//                  memory), sorts the survivors by depth along the cell's centre ray (ties by Gaussian index, so the list is
//                  a pure function of the frame whatever order the atomics of k1_bin_fill ran in), reserves its slice of the
//                  index array with one 64-bit atomic and writes it.  No count-then-write second pass, no separate sort
//                  kernel, no host round trip: capacities come from the previous frame (checked afterwards).
// Lists are (list_off[cell], list_cnt[cell]) pairs; their placement in the index array is arbitrary.
// ------------------------------------------------------------------------------------------------
constexpr int BIN_CX = 4, BIN_CY = 8;   // cells per bin
constexpr int LEAF_CAP = 512;           // entries a leaf warp sorts in shared memory (longer lists: global-memory path)
constexpr int LEAF_WARPS = 8;

struct BinGrid
{
    int nbx, nby;            // bins per axis
    // closed-form slab test per axis: |u A + B| <= lim |n(u)| with n(u) = u P + Q  (x: P = U x R, Q = U x Wv; y: P = R x U, Q = R x Wv)
    float px[3], qx[3], py[3], qy[3];
    float pp_x, pq_x, qq_x, pp_y, pq_y, qq_y; // |P|^2, P.Q, |Q|^2
    float u_min, u_max, v_min, v_max;          // plane coordinates of the first / last pixel sample of the image
};

// interval of plane coordinates u whose plane (through the apex, containing the other screen axis) comes within `lim` of p:
// |u a + b| <= lim |n(u)|.  |n(u)| is convex in u, so over any interval it is bounded by its values at the ends: a first,
// loose interval from the bound over the whole image, then one refinement with the bound over that interval.
__device__ __forceinline__ bool slab_interval(float a, float b, float lim, float pp, float pq, float qq, float umin, float umax, float &lo, float &hi)
{
    auto nlen = [&](float u) { return sqrtf(fmaxf(fmaf(u, fmaf(u, pp, 2.f * pq), qq), 0.f)); };
    float w = lim * fmaxf(nlen(umin), nlen(umax)) * 1.0001f;
    lo = umin;
    hi = umax;
    for (int it = 0; it < 2; ++it)
    {
        if (fabsf(a) <= 1e-30f)
        {
            if (fabsf(b) > w) return false; // parallel to every plane of this axis and too far
            lo = umin;
            hi = umax;
            return true;
        }
        const float r = 1.f / a;
        const float e0 = (-b - w) * r, e1 = (-b + w) * r;
        lo = fmaxf(fminf(e0, e1), umin);
        hi = fminf(fmaxf(e0, e1), umax);
        if (!(lo <= hi)) return false; // (also rejects NaN)
        w = lim * fmaxf(nlen(lo), nlen(hi)) * 1.0001f;
    }
    return true;
}

// pixel column (row) -> cell index along that axis (cells restart at every reference-tile edge when the tile size is not a
// multiple of the cell size)
__device__ __forceinline__ int cell_of_pixel(int x, int tile_px, int cells_per_tile, int cell_px, int uniform)
{
    if (uniform) return x / cell_px;
    const int t = x / tile_px;
    return t * cells_per_tile + (x - t * tile_px) / cell_px;
}

// bins [bx0, bx1] x [by0, by1] a Gaussian may be listed in (false: none)
__device__ __forceinline__ bool bin_rect(const FrameGeom &G, const BinGrid &B, const float4 a /* oc.xyz, sigma */, int &bx0, int &bx1, int &by0, int &by1)
{
    const float lim = G.bound_k * a.w + 1e-6f * (fabsf(a.x) + fabsf(a.y) + fabsf(a.z));
    if (!(lim >= 0.f) || !(fabsf(a.x) + fabsf(a.y) + fabsf(a.z) <= 3.0e38f)) return false; // NaN / inf records are listed nowhere (the exact test rejects them too)
    const float ax = a.x * B.px[0] + a.y * B.px[1] + a.z * B.px[2], bxq = a.x * B.qx[0] + a.y * B.qx[1] + a.z * B.qx[2];
    const float ay = a.x * B.py[0] + a.y * B.py[1] + a.z * B.py[2], byq = a.x * B.qy[0] + a.y * B.qy[1] + a.z * B.qy[2];
    float ulo, uhi, vlo, vhi;
    if (!slab_interval(ax, bxq, lim, B.pp_x, B.pq_x, B.qq_x, B.u_min, B.u_max, ulo, uhi)) return false;
    if (!slab_interval(ay, byq, lim, B.pp_y, B.pq_y, B.qq_y, B.v_min, B.v_max, vlo, vhi)) return false;
    // plane coordinate -> pixel (u = -1 + x / half_w), one pixel of slack for the rounding of the conversion
    const int xa = max(0, (int)floorf((ulo + 1.f) * G.half_w) - 1), xb = min(G.W - 1, (int)ceilf((uhi + 1.f) * G.half_w) + 1);
    const int ya = max(G.row_begin, (int)floorf((vlo + 1.f) * G.half_h) - 1), yb = min(G.row_end - 1, (int)ceilf((vhi + 1.f) * G.half_h) + 1);
    if (xa > xb || ya > yb) return false;
    bx0 = cell_of_pixel(xa, G.tile_w, G.cptx, CELL_W, G.uniform) / BIN_CX;
    bx1 = cell_of_pixel(xb, G.tile_w, G.cptx, CELL_W, G.uniform) / BIN_CX;
    by0 = cell_of_pixel(ya, G.tile_h, G.cpty, CELL_H, G.uniform) / BIN_CY;
    by1 = cell_of_pixel(yb, G.tile_h, G.cpty, CELL_H, G.uniform) / BIN_CY;
    return true;
}

// FILL = false: bin_count[b] += 1 per covered bin; FILL = true: append the Gaussian to every covered bin's list
template <bool FILL>
__global__ void __launch_bounds__(256) k1_bin(const FrameGeom G, const BinGrid B, const float4 *__restrict__ cullrec, uint32_t n, uint32_t *__restrict__ bin_count,
                                              const uint32_t *__restrict__ bin_off, uint32_t *__restrict__ bin_idx, uint64_t idx_cap,
                                              unsigned long long *__restrict__ total)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mine = 0;
    if (i < n)
    {
        const float4 a = cullrec[2 * i];
        int bx0, bx1, by0, by1;
        // (a Gaussian dropped by the reference projection, rt.cpp:38-41, is in no list)
        if (!(G.use_ref && cullrec[2 * i + 1].w == 0.f) && bin_rect(G, B, a, bx0, bx1, by0, by1))
        {
            mine = (unsigned long long)(bx1 - bx0 + 1) * (unsigned long long)(by1 - by0 + 1);
            for (int by = by0; by <= by1; ++by)
                for (int bx = bx0; bx <= bx1; ++bx)
                {
                    const uint32_t b = (uint32_t)(by * B.nbx + bx);
                    const uint32_t pos = atomicAdd(&bin_count[b], 1u);
                    if (FILL)
                    {
                        const uint64_t at = (uint64_t)bin_off[b] + pos;
                        if (at < idx_cap) bin_idx[at] = i;
                    }
                }
        }
    }
    if (!FILL)
    {
        // entries needed, in 64 bits (the scan of the bin counts is 32-bit)
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total, mine);
    }
}

// ---- per-warp sorts ----------------------------------------------------------------------------------------------------
// An entry is one 64-bit word: (order-preserving bits of the depth key) << 32 | Gaussian index, so "ascending by depth, ties
// by index" is a single unsigned comparison and the sorted order is a pure function of the frame.
__device__ __forceinline__ unsigned long long sort_word(float key, uint32_t val)
{
    const uint32_t b = __float_as_uint(key);
    const uint32_t o = b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u); // unsigned order = float order (keys are never NaN)
    return ((unsigned long long)o << 32) | val;
}

// n <= 32 M: rank of every element by counting (n^2 / 32 comparisons per lane, no barriers inside); writes out[rank] = index
template <int M>
__device__ __forceinline__ void warp_rank_sort(const unsigned long long *kv, uint32_t n, uint32_t *out, int lane)
{
    unsigned long long mine[M];
    uint32_t rank[M];
#pragma unroll
    for (int m = 0; m < M; ++m)
    {
        const uint32_t i = (uint32_t)lane + 32u * m;
        mine[m] = i < n ? kv[i] : 0xFFFFFFFFFFFFFFFFull;
        rank[m] = 0;
    }
#pragma unroll 4
    for (uint32_t j = 0; j < n; ++j)
    {
        const unsigned long long other = kv[j];
#pragma unroll
        for (int m = 0; m < M; ++m) rank[m] += other < mine[m] ? 1u : 0u;
    }
#pragma unroll
    for (int m = 0; m < M; ++m)
        if ((uint32_t)lane + 32u * m < n) out[rank[m]] = (uint32_t)mine[m];
}

// Bitonic sorting network in its all-ascending form (every merge starts with a "flip" step, partner i ^ (k - 1), then the
// half-cleaners i ^ j): every compare-exchange leaves the smaller word at the lower index, so the n real entries behave as if
// padded with +inf up to the next power of two without the padding ever being stored -- a partner at or beyond n is a no-op.
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long *kv, uint32_t n, int lane)
{
    uint32_t m = 2;
    while (m < n) m <<= 1;
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k; j > 1; j >>= 1)
        {
            // j == k: flip step; j < k: half-cleaner of distance j / 2
            for (uint32_t i = lane; i < n; i += 32)
            {
                const uint32_t l = (j == k) ? (i ^ (k - 1)) : (i ^ (j >> 1));
                if (l > i && l < n)
                {
                    const unsigned long long a = kv[i], b = kv[l];
                    if (b < a)
                    {
                        kv[i] = b;
                        kv[l] = a;
                    }
                }
            }
            __syncwarp();
        }
}

// depth key of Gaussian gi along the unit centre ray d of a cell; non-finite / huge depths are clamped below the padding key
__device__ __forceinline__ float depth_key(const float4 a, const float *d)
{
    const float k = a.x * d[0] + a.y * d[1] + a.z * d[2];
    return (k == k) ? fminf(fmaxf(k, -2.9e38f), 2.9e38f) : 2.9e38f;
}

// lists longer than LEAF_CAP: the same network on the global index array, keys refetched from the cull records
__device__ void warp_bitonic_sort_global(uint32_t *idx, uint32_t n, const float4 *__restrict__ cullrec, const float *d, int lane)
{
    uint32_t m = 2;
    while (m < n) m <<= 1;
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k; j > 1; j >>= 1)
        {
            for (uint32_t i = lane; i < n; i += 32)
            {
                const uint32_t l = (j == k) ? (i ^ (k - 1)) : (i ^ (j >> 1));
                if (l > i && l < n)
                {
                    const uint32_t vi = idx[i], vl = idx[l];
                    if (sort_word(depth_key(cullrec[2 * vl], d), vl) < sort_word(depth_key(cullrec[2 * vi], d), vi))
                    {
                        idx[i] = vl;
                        idx[l] = vi;
                    }
                }
            }
            __syncwarp();
        }
}

struct LeafArgs
{
    const float4 *cullrec;
    // SRC 0: the bin lists; SRC 1: the caller's per-tile record ranges (tile_off[t] .. tile_off[t + 1], no index array)
    const uint32_t *bin_off, *bin_count, *bin_idx;
    uint64_t bin_cap;
    const uint32_t *tile_off;
    int nbx;
    uint32_t *list_off, *list_cnt, *list_idx;
    unsigned long long *cursor; // entries reserved so far (64-bit: a wrap past 2^32 is seen, not silently reused)
    uint64_t idx_cap;
    uint32_t cell_begin, n_cells; // cells [cell_begin, n_cells) are launched: the cell rows of the rendered band
    uint8_t *list_wide; // per cell: 1 = a long list whose band is most of the list (see queue_key)
    float wide_frac;
};

template <int SRC>
__global__ void __launch_bounds__(LEAF_WARPS * 32) k1_leaf(const FrameGeom G, const LeafArgs L)
{
    __shared__ unsigned long long s_kv[LEAF_WARPS][LEAF_CAP];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t cell = L.cell_begin + blockIdx.x * LEAF_WARPS + w;
    if (cell >= L.n_cells) return;
    const int cx = (int)(cell % G.ncx), cy = (int)(cell / G.ncx);
    int x0, y0, cw, ch;
    cell_rect(G, cx, cy, x0, y0, cw, ch);
    uint32_t n = 0, off = 0;
    if (y0 + ch > G.row_begin && y0 < G.row_end) // cells outside the rendered row band get empty lists
    {
        uint32_t begin, end;
        if (SRC == 0)
        {
            const uint32_t b = (uint32_t)((cy / BIN_CY) * L.nbx + cx / BIN_CX);
            begin = L.bin_off[b];
            end = (uint32_t)min((uint64_t)begin + L.bin_count[b], L.bin_cap);
        }
        else
        {
            const uint32_t t = (uint32_t)((cy / G.cpty) * G.tiles_x + cx / G.cptx);
            begin = L.tile_off[t];
            end = L.tile_off[t + 1];
        }
        CullRect rc;
        make_rect(G, x0, x0 + cw, y0, y0 + ch, rc);
        // unit centre ray of the cell (the depth order K2's banded evaluation walks)
        const float u = -1.f + ((float)x0 + 0.5f * (float)(cw - 1)) / G.half_w, v = -1.f + ((float)y0 + 0.5f * (float)(ch - 1)) / G.half_h;
        float d[3];
        for (int i = 0; i < 3; ++i) d[i] = (G.inv0[i] * u + G.inv1[i] * v) + G.inv3[i] - G.origin[i];
        const float inv = rsqrtf(fmaxf(dot3(d, d), 1e-30f));
        for (int i = 0; i < 3; ++i) d[i] *= inv;
        unsigned long long *kv = s_kv[w];
        // the index of the NEXT step's candidate is fetched one step ahead (index -> record is a dependent pair of loads)
        uint32_t gi_next = (SRC == 0 && begin + lane < end) ? L.bin_idx[begin + lane] : begin + lane;
        for (uint32_t k = begin; k < end; k += 32)
        {
            const uint32_t e = k + lane;
            bool pass = false;
            const uint32_t gi = gi_next;
            if (SRC == 0) { if (e + 32 < end) gi_next = L.bin_idx[e + 32]; }
            else gi_next = e + 32;
            float kd = 0.f;
            if (e < end)
            {
                const float4 a = L.cullrec[2 * gi]; // (oc.xyz, sigma)
                const float4 cr = G.use_ref ? L.cullrec[2 * gi + 1] : make_float4(0.f, 0.f, 0.f, 1.f);
                pass = cull_test(G, rc, a, a.w, cr);
                kd = depth_key(a, d);
            }
            const uint32_t ballot = __ballot_sync(0xffffffffu, pass);
            const uint32_t pos = n + __popc(ballot & ((1u << lane) - 1u));
            if (pass && pos < (uint32_t)LEAF_CAP) kv[pos] = sort_word(kd, gi);
            n += __popc(ballot);
        }
        if (n)
        {
            unsigned long long at = 0;
            if (lane == 0) at = atomicAdd(L.cursor, (unsigned long long)n);
            at = __shfl_sync(0xffffffffu, at, 0);
            const bool fits = at + n <= L.idx_cap; // (an overflow is reported by the host after the frame; nothing is written)
            off = (uint32_t)at;
            __syncwarp();
            if (!fits) { /* counted, not stored */ }
            else if (n <= 32u) warp_rank_sort<1>(kv, n, L.list_idx + off, lane);
            else if (n <= 64u) warp_rank_sort<2>(kv, n, L.list_idx + off, lane);
            else if (n <= 96u) warp_rank_sort<3>(kv, n, L.list_idx + off, lane);
            else if (n <= 128u) warp_rank_sort<4>(kv, n, L.list_idx + off, lane);
            else if (n <= (uint32_t)LEAF_CAP)
            {
                warp_bitonic_sort(kv, n, lane);
                for (uint32_t i = lane; i < n; i += 32) L.list_idx[off + i] = (uint32_t)kv[i];
            }
            else
            {
                // longer than the shared-memory buffer: rescan straight into the index array, then sort it there
                uint32_t m = 0;
                for (uint32_t k = begin; k < end; k += 32)
                {
                    const uint32_t e = k + lane;
                    bool pass = false;
                    uint32_t gi = 0;
                    if (e < end)
                    {
                        gi = SRC == 0 ? L.bin_idx[e] : e;
                        const float4 a = L.cullrec[2 * gi];
                        const float4 cr = G.use_ref ? L.cullrec[2 * gi + 1] : make_float4(0.f, 0.f, 0.f, 1.f);
                        pass = cull_test(G, rc, a, a.w, cr);
                    }
                    const uint32_t ballot = __ballot_sync(0xffffffffu, pass);
                    if (pass) L.list_idx[off + m + __popc(ballot & ((1u << lane) - 1u))] = gi;
                    m += __popc(ballot);
                }
                __syncwarp();
                warp_bitonic_sort_global(L.list_idx + off, n, L.cullrec, d, lane);
            }
        }
    }
    // How wide is the band of a long list?  For its middle entry as the emitter (samples mu - 4 sigma .. mu), count the entries
    // whose erf argument is not saturated everywhere (|t| < 5.5 somewhere): in depth, centre within 5.5 sqrt2 sigma_j.  When that is
    // most of the list the banded kernels have nothing to skip and the cell is queued for k2_render's in-loop test instead.
    // A pure function of the cell's list: the bands of a frame take the same decision as the whole frame.
    uint32_t wide = 0;
    if (n > (uint32_t)WIN_CAP && n <= (uint32_t)LONG_CAP_KEY && (unsigned long long)off + n <= L.idx_cap)
    {
        __syncwarp();
        float d[3];
        {
            const float u = -1.f + ((float)x0 + 0.5f * (float)(cw - 1)) / G.half_w, v = -1.f + ((float)y0 + 0.5f * (float)(ch - 1)) / G.half_h;
            for (int i = 0; i < 3; ++i) d[i] = (G.inv0[i] * u + G.inv1[i] * v) + G.inv3[i] - G.origin[i];
            const float inv = rsqrtf(fmaxf(dot3(d, d), 1e-30f));
            for (int i = 0; i < 3; ++i) d[i] *= inv;
        }
        const float4 am = L.cullrec[2 * L.list_idx[off + n / 2]];
        const float hi = depth_key(am, d), lo = hi - 4.f * am.w;
        uint32_t open = 0;
        for (uint32_t i0 = 0; i0 < n; i0 += 32)
        {
            const uint32_t i = i0 + lane;
            bool o = false;
            if (i < n)
            {
                const float4 a = L.cullrec[2 * L.list_idx[off + i]];
                const float k = depth_key(a, d), half = 7.7782f * a.w;
                o = k + half > lo && k - half < hi;
            }
            open += __popc(__ballot_sync(0xffffffffu, o));
        }
        wide = (float)open > L.wide_frac * (float)n ? 1u : 0u;
    }
    if (lane == 0)
    {
        L.list_off[cell] = off;
        L.list_cnt[cell] = n;
        if (L.list_wide != nullptr) L.list_wide[cell] = (uint8_t)wide;
    }
}
