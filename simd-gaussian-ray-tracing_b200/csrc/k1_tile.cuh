// k1_tile.cuh -- K0 (per-Gaussian frame records), the culling predicate, the literal per-tile lists, scans, cost histogram and work queue
// (the bounded per-cell lists are built by k1_bin.cuh).
// A fragment of vrt_cuda.cu: included there, in this order, inside its anonymous namespace.  Not a stand-alone header.
#pragma once

// ------------------------------------------------------------------------------------------------
// K0: per-Gaussian frame constants
// ------------------------------------------------------------------------------------------------
// cull record, 32 B: (oc.xyz, sigma) and (mu'.x, mu'.y, 3.3 sigma', valid) of the reference's tiling projection (rt.cpp:35-45)
// scene_info[0] = max |albedo component| as float bits (all non-negative, so the unsigned order is the float order),
// scene_info[1] = 1 when some magnitude is negative or some sigma is not positive: T(s) is then not monotone and K2 must not
// terminate early.
__global__ void k0_prepare(const FrameGeom G, const float *__restrict__ aos, uint64_t n, Rec *__restrict__ rec, float4 *__restrict__ cullrec,
                           uint32_t *__restrict__ scene_info)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *g = aos + i * 10;
    const float ax = g[0], ay = g[1], az = g[2], aw = g[3];
    const float mx = g[4], my = g[5], mz = g[6], mw = g[7];
    const float sigma = g[8], mag = g[9];
    Rec r;
    r.a = make_float4(mx - G.origin[0], my - G.origin[1], mz - G.origin[2], mw * mw);
    const float rr = 1.f / (1.41421356237309504880f * sigma);
    r.b = make_float4(rr, LOG2E / (2.f * sigma * sigma), sigma * mag * SQRT_PI_2 * LOG2E, sigma);
    r.c = make_float4(ax, ay, az, aw);
    rec[i] = r;
    if (cullrec != nullptr)
    {
        // proj = view * (mu.xyz, 1), GLM operand order (c0 x + c1 y) + (c2 z + c3 w)
        const float *v = G.view;
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
    if (scene_info != nullptr)
    {
        // the running maximum is read first: after the first few Gaussians almost no thread issues an atomic
        float am = fmaxf(fmaxf(fabsf(ax), fabsf(ay)), fmaxf(fabsf(az), fabsf(aw)));
        am = (am == am) ? am : 3.0e38f;
        const uint32_t bits = __float_as_uint(am);
        if (bits > scene_info[0]) atomicMax(&scene_info[0], bits);
        if ((mag < 0.f || !(sigma > 0.f)) && scene_info[1] == 0u) atomicOr(&scene_info[1], 1u);
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
__device__ __forceinline__ void make_rect(const FrameGeom &G, int x0, int x1, int y0, int y1, CullRect &rc)
{
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
__device__ __forceinline__ bool near_frustum(const FrameGeom &G, const CullRect &rc, const float *p, float dl, float dr, float db, float dt, float sgn, float lim)
{
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

__device__ __forceinline__ bool cull_test(const FrameGeom &G, const CullRect &rc, const float4 a, const float sigma, const float4 cr)
{
    if (G.use_ref)
    {
        if (cr.w == 0.f) return false;
        const float hx = G.tw / 2, hy = G.th / 2;
        if (rc.exact_tile)
        {
            if (!(ref_axis(G.tile_cx[rc.tx0], cr.x, hx, cr.z) && ref_axis(G.tile_cy[rc.ty0], cr.y, hy, cr.z))) return false;
        }
        else
        {
            // |c - mu| - |c| is monotone in c, so a tile range passes iff one of its end tiles does;
            // the slack keeps the coarse level conservative against rounding of the exact test.
            const float s = cr.z + 1e-4f;
            const bool px = ref_axis(G.tile_cx[rc.tx0], cr.x, hx, s) || ref_axis(G.tile_cx[rc.tx1], cr.x, hx, s);
            const bool py = ref_axis(G.tile_cy[rc.ty0], cr.y, hy, s) || ref_axis(G.tile_cy[rc.ty1], cr.y, hy, s);
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
        if (!(near_frustum(G, rc, p, dl, dr, db, dt, 1.f, lim) || near_frustum(G, rc, p, dl, dr, db, dt, -1.f, lim))) return false;
    }
    return true;
}

// pure REFERENCE lists (one list per reference tile), level 1 with children = tiles
template <bool WRITE>
__global__ void __launch_bounds__(256) k1_cull_tiles(const FrameGeom G, const Rec *__restrict__ rec, const float4 *__restrict__ cullrec, uint32_t n_root,
                                                     uint32_t *__restrict__ counts, const uint32_t *__restrict__ offsets,
                                                     uint32_t *__restrict__ out_idx, uint32_t n_tiles)
{
    // one CTA per tile; ordered compaction across the CTA's 8 warps
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    const uint32_t tile = blockIdx.x;
    if (tile >= n_tiles) return;
    const int tx = tile % G.tiles_x, ty = tile / G.tiles_x;
    const float cx = G.tile_cx[tx], cy = G.tile_cy[ty];
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

__device__ __forceinline__ void cell_rect(const FrameGeom &G, int cx, int cy, int &x0, int &y0, int &w, int &h);

// exclusive scan of n counts into n+1 offsets; single CTA of 1024 threads (n <= a few million).  The array is walked in
// chunks of 1024 consecutive elements (coalesced), each chunk scanned by warp shuffles + one shared-memory pass over the 32
// warp totals, with a running carry: two barriers per chunk (a Hillis-Steele scan over per-thread sub-arrays needed twenty
// and read its input with a stride: 20 us for the 16 K bin counts of a 4096^2 frame, 5 us now).
__global__ void __launch_bounds__(1024) k1_scan(const uint32_t *__restrict__ counts, uint32_t *__restrict__ offsets, uint32_t n)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024u)
    {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? counts[i] : 0u;
        uint32_t inc = v;
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) s_warp[w] = inc;
        const uint32_t carry = s_carry;
        __syncthreads();
        if (w == 0)
        {
            const uint32_t tot = s_warp[lane];
            uint32_t ws = tot;
            for (int d = 1; d < 32; d <<= 1)
            {
                const uint32_t t = __shfl_up_sync(0xffffffffu, ws, d);
                if (lane >= d) ws += t;
            }
            s_warp[lane] = ws - tot; // exclusive prefix of the warp totals
            if (lane == 31) s_carry = carry + ws;
        }
        __syncthreads();
        if (i < n) offsets[i] = carry + s_warp[w] + inc - v;
    }
    __syncthreads();
    if (threadIdx.x == 0) offsets[n] = s_carry;
}

// sum of n 32-bit counts in 64 bits (the exclusive scans produce 32-bit offsets: a grand total past 2^32 must be seen, not wrapped)
__global__ void __launch_bounds__(256) k1_total64(const uint32_t *__restrict__ counts, uint32_t n, unsigned long long *__restrict__ total)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = i < n ? counts[i] : 0u;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(total, v);
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
    unsigned long long terms_term; // K2b: terms dropped by the transmittance early exit
    unsigned long long leaf_entries; // k1_leaf: list entries reserved (the bump cursor; may exceed the index array's capacity)
    unsigned long long bin_entries;  // k1_bin: bin list entries needed
    int slice;                       // emitters per work item of split cells chosen for this frame (k1_pick_slice)
    int pad;
    unsigned long long n_huge;       // queued items whose list is longer than k2_band_long's cache (they lead the queue)
};

__device__ __forceinline__ uint32_t cell_list_id(const FrameGeom &G, int cx, int cy)
{
    if (G.list_kind == 0) return (uint32_t)(cy * G.ncx + cx);
    if (G.list_kind == 1) return (uint32_t)((cy / G.cpty) * G.tiles_x + (cx / G.cptx));
    return 0u;
}

__device__ __forceinline__ void cell_rect(const FrameGeom &G, int cx, int cy, int &x0, int &y0, int &w, int &h)
{
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

// emitters per work item of split cells: fixed by the host (a pinned slice / after the tile call returned) or, inside the
// tile call, chosen on the device from the frame's own statistics (k1_pick_slice) so that no host round trip is needed
__device__ __forceinline__ uint32_t frame_slice(const FrameGeom &G) { return G.slice > 0 ? (uint32_t)G.slice : (uint32_t)*G.slice_dev; }

// number of work items of a cell with an n-entry list
__device__ __forceinline__ uint32_t cell_items(const FrameGeom &G, uint32_t n, uint32_t cell)
{
    const uint32_t slice = frame_slice(G);
    if (n <= 3u * slice || cell >= (1u << ITEM_CELL_BITS)) return 1u;
    const uint32_t k = (n + slice - 1) / slice;
    return k <= (1u << (32 - ITEM_CELL_BITS)) ? k : 1u;
}

constexpr uint32_t HIST_KEYS = 8192; // key = min(n, HIST_KEYS - 1) (queue_key); the queue is filled in descending key order
// Queue key of a list (the queue is filled in descending key order): its length -- except that a long list K1 marked as WIDE
// (its band is most of the list: nothing for k2_band_long to gain) is keyed beyond LONG_CAP, with the lists that do not fit
// that kernel's cache, so that the head of the queue is exactly what k2_render's in-loop test works off.
constexpr uint32_t WIDE_KEY_SHIFT = 4096;
__device__ __forceinline__ uint32_t queue_key(uint32_t n, const uint8_t *__restrict__ wide, uint32_t id)
{
    if (wide != nullptr && n > (uint32_t)WIN_CAP && n <= (uint32_t)LONG_CAP_KEY && wide[id]) return n + WIDE_KEY_SHIFT;
    return min(n, HIST_KEYS - 1u);
}

// COUNT_ITEMS = false: listed terms and per-row cost; true: work items per list-length key (needs the frame's slice size)
template <bool COUNT_ITEMS>
__global__ void k1_hist(const FrameGeom G, const uint32_t *__restrict__ list_cnt, uint32_t *__restrict__ hist, TileStats *__restrict__ stats,
                        double *__restrict__ row_cost, int cy_begin, int cy_end, const uint8_t *__restrict__ wide)
{
    const int ncells = (cy_end - cy_begin) * G.ncx;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double terms = 0.0;
    uint32_t items = 0, split = 0;
    if (i < ncells)
    {
        const int cx = i % G.ncx, cy = cy_begin + i / G.ncx;
        int x0, y0, w, h;
        cell_rect(G, cx, cy, x0, y0, w, h);
        const int ya = max(y0, G.row_begin), yb = min(y0 + h, G.row_end);
        const uint32_t id = cell_list_id(G, cx, cy);
        const uint32_t n = list_cnt[id];
        if (yb > ya)
        {
            if (COUNT_ITEMS)
            {
                items = cell_items(G, n, (uint32_t)(cy * G.ncx + cx));
                atomicAdd(&hist[queue_key(n, wide, id)], items);
                split = items > 1 ? items : 0u;
            }
            else
            {
                terms = 5.0 * (double)n * (double)n * (double)(w * (yb - ya));
                if (row_cost != nullptr) atomicAdd(&row_cost[cy], terms);
            }
        }
    }
    // one atomic per warp for the frame totals (a per-thread atomic on one address serialises the whole launch)
    if (COUNT_ITEMS)
    {
        for (int o = 16; o > 0; o >>= 1)
        {
            items += __shfl_xor_sync(0xffffffffu, items, o);
            split += __shfl_xor_sync(0xffffffffu, split, o);
        }
        if ((threadIdx.x & 31) == 0 && items) atomicAdd(&stats->n_items, (unsigned long long)items);
        if ((threadIdx.x & 31) == 0 && split) atomicAdd(&stats->n_split, (unsigned long long)split);
    }
    else
    {
        for (int o = 16; o > 0; o >>= 1) terms += __shfl_xor_sync(0xffffffffu, terms, o);
        if ((threadIdx.x & 31) == 0 && terms != 0.0) atomicAdd(&stats->terms_listed, terms);
    }
}

__global__ void k1_list_stats(const uint32_t *__restrict__ list_cnt, uint32_t n_lists, TileStats *__restrict__ stats)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = 0;
    if (i < n_lists) n = list_cnt[i];
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

// Slice size of split cells, chosen on the device (the host-side twin is auto_slice in vrt_cuda.cu): as large as possible
// (every item repeats pass A), but no item may exceed a quarter of the average work per resident warp, or the longest lists
// would decide the frame time on small frames.
__global__ void k1_pick_slice(TileStats *__restrict__ stats, int *__restrict__ slice_dev, int pinned_slice, double warps)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int slice = pinned_slice;
    if (slice <= 0)
    {
        const double per_warp = stats->terms_listed / warps;
        slice = SLICE_MAX;
        while (slice > SLICE_MIN && 160.0 * slice * (double)stats->max_list > per_warp / 4.0) slice /= 2;
    }
    *slice_dev = slice;
    stats->slice = slice;
}

// hist (ascending key) -> start position of each key in a DESCENDING ordering; single CTA.  Also leaves, in stats->n_big, the
// number of items whose list is longer than the banded kernel's cache (they lead the queue).
__global__ void __launch_bounds__(1024) k1_hist_scan(uint32_t *__restrict__ hist, TileStats *__restrict__ stats)
{
    __shared__ uint32_t s_part[1024];
    constexpr int PER = HIST_KEYS / 1024;
    // thread t owns keys [t*PER, t*PER+PER) counted from the top: descending order
    const int t = threadIdx.x;
    uint32_t c[PER];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k)
    {
        c[k] = hist[HIST_KEYS - 1 - (t * PER + k)];
        sum += c[k];
    }
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
#pragma unroll
    for (int k = 0; k < PER; ++k)
    {
        const int key = HIST_KEYS - 1 - (t * PER + k);
        hist[key] = run;
        if (key == WIN_CAP) stats->n_big = run; // items with a longer list start the queue
        if (key == LONG_CAP_KEY) stats->n_huge = run;
        run += c[k];
    }
}

__global__ void k1_order(const FrameGeom G, const uint32_t *__restrict__ list_cnt, uint32_t *__restrict__ cursor, uint32_t *__restrict__ queue, uint32_t *__restrict__ cell_slot,
                         uint32_t *__restrict__ split_cursor, int cy_begin, int cy_end, uint32_t queue_cap, const uint8_t *__restrict__ wide)
{
    const int ncells = (cy_end - cy_begin) * G.ncx;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncells) return;
    const int cx = i % G.ncx, cy = cy_begin + i / G.ncx;
    int x0, y0, w, h;
    cell_rect(G, cx, cy, x0, y0, w, h);
    if (min(y0 + h, G.row_end) <= max(y0, G.row_begin)) return;
    const uint32_t id = cell_list_id(G, cx, cy);
    const uint32_t n = list_cnt[id];
    const uint32_t cell = (uint32_t)(cy * G.ncx + cx);
    const uint32_t items = cell_items(G, n, cell);
    const uint32_t pos = atomicAdd(&cursor[queue_key(n, wide, id)], items);
    for (uint32_t k = 0; k < items; ++k)
        if (pos + k < queue_cap) queue[pos + k] = cell | (k << ITEM_CELL_BITS);
    // split cells get `items` consecutive slots of the partial-radiance buffer (the slot order is irrelevant: K3' sums a
    // cell's slices in slice order)
    cell_slot[cell] = items > 1 ? atomicAdd(split_cursor, items) : NO_SLOT;
}
