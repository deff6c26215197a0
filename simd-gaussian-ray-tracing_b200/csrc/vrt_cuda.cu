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
// Layout: vrt_common.cuh (constants, records, erf) / k1_tile.cuh (K0, K1) / k2_render.cuh (K2, K3) / k2_band.cuh (K2b, the
// default kernel on depth-sorted lists) / k2_variant.cuh / probes.cuh are fragments of this translation unit; this file holds
// the context, the launch logic and the C ABI.  Nothing about a frame lives in per-device state (no __constant__ geometry):
// contexts are independent even on one GPU.
//
// No CPU fallback: every entry point fails without a CUDA device.  Nothing under oracle/ is used here.
#include "vrt_cuda.h"
#include "vrt_approx_tables.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

// ------------------------------------------------------------------------------------------------
// device code: one translation unit, split by kernel family
// ------------------------------------------------------------------------------------------------
namespace
{
#include "vrt_common.cuh"
#include "k1_tile.cuh"
#include "k1_bin.cuh"
#include "k2_render.cuh"
#include "k2_variant.cuh"
#include "k2_band.cuh"
#include "k2_long.cuh"
#include "probes.cuh"

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
    DevBuf bin_count, bin_off, bin_idx;   // per-Gaussian screen binning (k1_bin)
    DevBuf ccounts, coffsets, cidx;       // the lists K2 walks: entries [coffsets[l], coffsets[l] + ccounts[l]) of cidx
    DevBuf slice_dev;                     // slice size chosen on the device during the tile call
    DevBuf hist, queue, stats, counter, rowcost, scan_tmp, cell_slot, partial, cwide;
    DevBuf tile_centres; // 2 x 1024 floats: the reference's float-accumulated tile centres of the current frame
    DevBuf scene_info;   // K0: max |albedo| bits, non-monotone flag
    // interruption: the persistent render warps poll a word in DEVICE memory (an L2 hit per work item, not a PCIe read);
    // vrt_cuda_abort writes it from any thread with a 4-byte copy on a stream of its own, which overlaps the running kernel
    uint32_t *abort_host = nullptr; // pinned staging word
    uint32_t *abort_dev = nullptr;
    cudaStream_t abort_stream = nullptr;
    // caller buffers this context page-locked (cudaHostRegister) so that the copies of vrt_cuda_set_gaussians /
    // vrt_cuda_render run at PCIe speed instead of through the driver's staging buffer: the reference's callers hand in
    // plain aligned_malloc memory (main.cpp:245) and reuse it every frame
    struct Pinned { void *ptr = nullptr; size_t bytes = 0; };
    Pinned pinned[4];
    int pinned_next = 0;
    int pin_mode = 0; // opt-in (vrt_cuda_set_host_pinning): the caller promises its buffers outlive the registration
    DevBuf out_image, out_rad;
    // Pageable caller buffers (what the reference's main hands in, main.cpp:245) without the pinning opt-in: the library keeps
    // page-locked staging of its own.  The frame is written into `stage_out` by K3 (mapped: 16-byte stores over PCIe while the
    // frame is computed) and copied to the caller's image by a few host threads; the scene is copied into `stage_in` by the same
    // threads, chunk by chunk, each chunk's upload overlapping the next chunk's host copy.  The driver's own staging of a
    // pageable cudaMemcpy runs at a third of that rate.
    struct HostStage { void *p = nullptr; void *dev = nullptr; size_t cap = 0; };
    HostStage stage_out, stage_in, stage_rad;
    cudaEvent_t ev_stage = nullptr; // completion of the last upload out of stage_in
    struct PeerImage { void *ptr = nullptr; bool owned = false; };
    std::vector<PeerImage> peer_images; // vrt_cuda_peer_image_create / _open
    DevBuf tile_aos; // host-supplied tile lists (concatenated)
    DevBuf tile_off;
    DevBuf lit_offsets, lit_counts, lit_idx; // literal lists (tile_gaussians membership) kept for vrt_cuda_get_lists while K2 uses the visible lists

    // state of the last tile()
    bool have_lists = false;
    bool lists_from_host = false;
    bool lists_sorted = false;
    uint32_t n_big = 0, n_huge = 0, n_split = 0;
    float long_wide = 0.4f; // K1 marks a long list as wide (-> k2_render<WIN>) when its middle emitter must evaluate more than this share of it (VRT_CUDA_LONG_WIDE)
    int win_minb = 3; // resident CTAs per SM of k2_render<WIN>: 3 (4-warp CTAs, 12 warps/SM, 168 registers; 5-7 % faster on the OBJ scenes, same bits) or 1 (8 warps, 236 registers: round 1); VRT_CUDA_WIN_MINB
    bool long_band = true; // lists beyond k2_band's cache go to k2_band_long (VRT_CUDA_LONG_BAND=0: to k2_render's in-loop test, as in round 1)
    FrameGeom geom{};
    uint32_t n_lists = 0;
    uint64_t n_entries = 0;
    uint32_t n_queue = 0, queue_cap = 0;
    int cy_begin = 0, cy_end = 0;
    uint64_t bin_hint = 0, idx_hint = 0; // capacities (entries) that sufficed for the previous frame
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
    double q_terms_listed = 0.0; // listed terms / longest list of the lists the queue was last built for
    uint32_t q_max_list = 0;
    int tune_slice = 0; // 0 = automatic (auto_slice)
    int tune_q = 0; // 0 = automatic: 8 emitters per register block, 4 when the lists are short
    int tune_pack = 1;
    int tune_band_ctas = 4; // resident CTAs per SM of the banded kernel (4 or 5)
    std::vector<float> centres_host;
    uint32_t tiled_list_mode = 0; // what the current lists were built for (vrt_cuda_render_device checks the frame against it)
    int tiled_tiles_x = 0, tiled_tiles_y = 0;
    float tiled_bound = 0.f;
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
    // a work item packs the cell id into ITEM_CELL_BITS bits (k1_order / k2_render)
    if ((uint64_t)G.ncx * (uint64_t)G.ncy > (1ull << ITEM_CELL_BITS))
        return fail(ctx, VRT_CUDA_E_INVALID, "image %ux%u has more than %llu 8x4-pixel cells (about 134 Mpixel): render it in row bands of separate frames", f->width, f->height,
                    1ull << ITEM_CELL_BITS);
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

// The tile centres go to the context's own device buffer (stream-ordered after the previous frame's kernels); the geometry
// itself is a kernel argument.
int upload_geom(vrt_cuda_ctx *ctx, FrameGeom &G, const std::vector<float> &cxs, const std::vector<float> &cys)
{
    if (int rc = reserve(ctx, ctx->tile_centres, sizeof(float) * 2048)) return rc;
    ctx->centres_host.resize(2048);
    std::memcpy(ctx->centres_host.data(), cxs.data(), sizeof(float) * 1024);
    std::memcpy(ctx->centres_host.data() + 1024, cys.data(), sizeof(float) * 1024);
    CU(cudaMemcpyAsync(ctx->tile_centres.p, ctx->centres_host.data(), sizeof(float) * 2048, cudaMemcpyHostToDevice, ctx->stream));
    G.tile_cx = (const float *)ctx->tile_centres.p;
    G.tile_cy = G.tile_cx + 1024;
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
// Slice size of split cells: as large as possible (every item repeats pass A), but no item may exceed a quarter of the
// average work per resident warp, or the longest lists would decide the frame time on small frames.  `share` scales the
// listed work (a rank that renders 1/N of the frame the lists were built for).
int auto_slice(const vrt_cuda_ctx *ctx, double share)
{
    const double per_warp = ctx->q_terms_listed * share / ((double)ctx->sm_count * 12.0);
    int slice = SLICE_MAX;
    while (slice > SLICE_MIN && 160.0 * slice * (double)ctx->q_max_list > per_warp / 4.0) slice /= 2;
    return slice;
}

// Statistics, slice choice, cost histogram and the descending-cost work queue of the band's cells.  Everything is enqueued
// on the context's stream; nothing is read back here.  queue_cap_hint > 0: the queue is sized from that bound (bounded per-cell
// lists: cells + entries / SLICE_MIN) and there is no host round trip at all; 0: the item count is read back once (literal
// list modes, whose per-cell item counts have no useful a-priori bound).  The caller reads the TileStats after its own final
// synchronisation (finish_tile).
int build_queue(vrt_cuda_ctx *ctx, uint64_t queue_cap_hint)
{
    FrameGeom &G = ctx->geom;
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
    const int ncells = std::max(0, cye - cyb) * G.ncx;
    if (int rc = reserve(ctx, ctx->hist, sizeof(uint32_t) * HIST_KEYS)) return rc;
    if (int rc = reserve(ctx, ctx->counter, sizeof(uint32_t) * 8)) return rc;
    if (int rc = reserve(ctx, ctx->rowcost, sizeof(double) * (size_t)G.ncy)) return rc;
    if (int rc = reserve(ctx, ctx->slice_dev, sizeof(int))) return rc;
    if (int rc = reserve(ctx, ctx->cell_slot, sizeof(uint32_t) * (size_t)G.ncx * G.ncy)) return rc;
    CU(cudaMemsetAsync(ctx->hist.p, 0, sizeof(uint32_t) * HIST_KEYS, ctx->stream));
    CU(cudaMemsetAsync(ctx->rowcost.p, 0, sizeof(double) * (size_t)G.ncy, ctx->stream));
    CU(cudaMemsetAsync((uint32_t *)ctx->counter.p + 2, 0, sizeof(uint32_t), ctx->stream));
    // the slice size is chosen on the device (k1_pick_slice) and read through G.slice_dev until the tile call has returned
    G.slice = 0;
    G.slice_dev = (const int *)ctx->slice_dev.p;
    TileStats *stats = (TileStats *)ctx->stats.p;
    const uint32_t *lcnt = (const uint32_t *)ctx->ccounts.p;
    const uint8_t *wide = (ctx->lists_sorted && G.list_kind == 0 && ctx->long_band) ? (const uint8_t *)ctx->cwide.p : nullptr;
    const int tb = 256, gb = std::max(1, (ncells + tb - 1) / tb);
    k1_list_stats<<<(ctx->n_lists + 255) / 256, 256, 0, ctx->stream>>>(lcnt, ctx->n_lists, stats);
    k1_hist<false><<<gb, tb, 0, ctx->stream>>>(G, lcnt, (uint32_t *)ctx->hist.p, stats, (double *)ctx->rowcost.p, cyb, cye, nullptr);
    k1_pick_slice<<<1, 32, 0, ctx->stream>>>(stats, (int *)ctx->slice_dev.p, ctx->tune_slice, (double)ctx->sm_count * 12.0);
    k1_hist<true><<<gb, tb, 0, ctx->stream>>>(G, lcnt, (uint32_t *)ctx->hist.p, stats, nullptr, cyb, cye, wide);
    k1_hist_scan<<<1, 1024, 0, ctx->stream>>>((uint32_t *)ctx->hist.p, stats);
    uint64_t cap = queue_cap_hint;
    if (cap == 0)
    {
        TileStats ts;
        CU(cudaMemcpyAsync(&ts, ctx->stats.p, sizeof(ts), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        cap = ts.n_items;
    }
    if (cap > 0xFFFFFFF0ull) return fail(ctx, VRT_CUDA_E_NOMEM, "work queue of %llu items", (unsigned long long)cap);
    if (int rc = reserve(ctx, ctx->queue, sizeof(uint32_t) * (size_t)std::max<uint64_t>(cap, 1))) return rc;
    ctx->queue_cap = (uint32_t)cap;
    k1_order<<<gb, tb, 0, ctx->stream>>>(G, lcnt, (uint32_t *)ctx->hist.p, (uint32_t *)ctx->queue.p, (uint32_t *)ctx->cell_slot.p, (uint32_t *)ctx->counter.p + 2, cyb, cye,
                                       (uint32_t)cap, wide);
    ctx->launches += 6;
    CU(cudaGetLastError());
    return 0;
}

// End of a tile call: the one host synchronisation.  Reads the frame's statistics (queue length, split cells, slice, the
// entry counts the capacity checks need) and fixes the slice size in the host copy of the geometry.
int finish_tile(vrt_cuda_ctx *ctx, TileStats &ts)
{
    CU(cudaMemcpyAsync(&ts, ctx->stats.p, sizeof(ts), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->q_terms_listed = ts.terms_listed;
    ctx->q_max_list = (uint32_t)ts.max_list;
    ctx->n_queue = (uint32_t)std::min<unsigned long long>(ts.n_items, ctx->queue_cap);
    ctx->n_split = (uint32_t)ts.n_split;
    ctx->n_big = (uint32_t)ts.n_big;
    ctx->n_huge = (uint32_t)ts.n_huge;
    ctx->geom.slice = ts.slice > 0 ? ts.slice : SLICE_MAX;
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
    // mean list length over the lists that are actually rendered: a row band leaves the per-cell lists outside it empty
    const uint64_t lists_in_band = ctx->geom.list_kind == 0 ? (uint64_t)std::max(0, ctx->cy_end - ctx->cy_begin) * ctx->geom.ncx : ctx->n_lists;
    // A pinned slice size (vrt_cuda_set_slice) says the caller composes one frame from bands rendered separately and wants
    // them bit-identical to the whole frame: the emitter block must then not depend on the band's own lists either (a band of
    // mean length 24.4 inside a frame of mean 23.6 once rendered with Q = 8 next to a Q = 4 frame: same sums, another grouping).
    const int q = ctx->tune_q ? ctx->tune_q : ((ctx->tune_slice || (lists_in_band && ctx->n_entries / lists_in_band >= 24)) ? 8 : 4);
    const bool p = ctx->tune_pack != 0;
    if (a.window)
    {
        // banded evaluation (depth-sorted index lists).  The queue is in descending list length: the items whose list does not
        // fit k2_band's per-warp cache lead it.  Those go to k2_band_long (one CTA per item, the list cached once per CTA in
        // dynamic shared memory sized to the frame's longest list), lists beyond ITS capacity and the
        // long lists K1 marked as wide (queued with them) to k2_render's in-loop saturation test, the rest to k2_band.
        static_assert(LONG_CAP == LONG_CAP_KEY, "k1_hist_scan marks the queue position of the lists beyond k2_band_long's cache");
        ctx->render_launches = 0;
        const uint32_t n_big = std::min(ctx->n_big, a.n_queue);
        const uint32_t n_huge = ctx->long_band ? std::min(ctx->n_huge, n_big) : n_big;
        if (n_huge)
        {
            RenderArgs big = a;
            big.n_queue = n_huge;
            if (ctx->win_minb == 3) launch_k2c<ERF, 8, true, 3, false, true>(ctx, big);
            else launch_k2c<ERF, 8, true, 1, false, true>(ctx, big);
            ctx->render_launches++;
        }
        if (n_big > n_huge)
        {
            const uint32_t cap = std::min<uint32_t>((uint32_t)LONG_CAP, (std::max<uint32_t>(ctx->q_max_list, WIN_CAP + 1) + 7u) & ~7u);
            const size_t smem = (size_t)cap * LONG_ENTRY_BYTES;
            int per_sm = 1;
            // 4 CTAs per SM (128 registers) while four caches fit, else 3 (168 registers)
            const bool four = 4 * (smem + 5120) <= 227u * 1024u;
            // (the opt-in ceiling is a per-function, per-device setting shared by every context: always the kernel's maximum, so that
            // two contexts with different list lengths cannot lower it under each other's launches)
            constexpr int long_smem_max = LONG_CAP * LONG_ENTRY_BYTES;
            if (four)
            {
                CU(cudaFuncSetAttribute(k2_band_long<ERF, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, long_smem_max));
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2_band_long<ERF, 4>, LONG_WARPS * 32, smem);
            }
            else
            {
                CU(cudaFuncSetAttribute(k2_band_long<ERF, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, long_smem_max));
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2_band_long<ERF, 3>, LONG_WARPS * 32, smem);
            }
            if (per_sm < 1) per_sm = 1;
            const uint32_t grid = std::max(1u, std::min(n_big - n_huge, (uint32_t)(ctx->sm_count * per_sm)));
            if (four) k2_band_long<ERF, 4><<<grid, LONG_WARPS * 32, smem, ctx->stream>>>(a, n_huge, n_big, cap);
            else k2_band_long<ERF, 3><<<grid, LONG_WARPS * 32, smem, ctx->stream>>>(a, n_huge, n_big, cap);
            ctx->render_launches++;
        }
        if (a.n_queue > n_big)
        {
            // 4 CTAs of 4 warps per SM (128 registers); 5 (96 registers, a few spills) is kept for comparison (vrt_cuda_set_band_tuning)
            int per_sm = 1;
            if (ctx->tune_band_ctas == 5) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2_band<ERF, 5>, BAND_WARPS * 32, 0);
            else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2_band<ERF, 4>, BAND_WARPS * 32, 0);
            if (per_sm < 1) per_sm = 1;
            const uint32_t want = (a.n_queue - n_big + BAND_WARPS - 1) / BAND_WARPS;
            const uint32_t grid = std::max(1u, std::min(want, (uint32_t)(ctx->sm_count * per_sm)));
            if (ctx->tune_band_ctas == 5) k2_band<ERF, 5><<<grid, BAND_WARPS * 32, 0, ctx->stream>>>(a, n_big);
            else k2_band<ERF, 4><<<grid, BAND_WARPS * 32, 0, ctx->stream>>>(a, n_big);
            ctx->render_launches++;
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

template <int FN>
void launch_approx_rate(vrt_cuda_ctx *ctx, int blocks, int iters)
{
    k_approx_rate<FN><<<blocks, 256, 0, ctx->stream>>>((float *)ctx->counter.p, iters);
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
    if (const char *e = std::getenv("VRT_CUDA_LONG_WIDE")) c->long_wide = (float)std::atof(e);
    if (const char *e = std::getenv("VRT_CUDA_WIN_MINB")) c->win_minb = std::atoi(e) == 1 ? 1 : 3;
    if (const char *e = std::getenv("VRT_CUDA_LONG_BAND")) c->long_band = std::atoi(e) != 0; // A/B knob: 0 sends long lists to k2_render<WIN>
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
    // the word the persistent render warps poll (see vrt_cuda_abort)
    if (cudaHostAlloc((void **)&c->abort_host, sizeof(uint32_t), cudaHostAllocDefault) != cudaSuccess ||
        cudaMalloc((void **)&c->abort_dev, sizeof(uint32_t)) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->abort_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMemset(c->abort_dev, 0, sizeof(uint32_t)) != cudaSuccess)
    {
        g_create_error = "cannot allocate the abort word";
        vrt_cuda_destroy(c);
        return VRT_CUDA_E_CUDA;
    }
    *c->abort_host = 0u;
    if (cudaEventCreateWithFlags(&c->ev_stage, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(c->ev_stage, c->stream) != cudaSuccess)
    {
        cudaGetLastError();
        c->ev_stage = nullptr; // (without it pageable uploads take the driver's staged copy)
    }
    *ctx_out = c;
    return 0;
}

void vrt_cuda_destroy(vrt_cuda_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &r : ctx->pinned)
        if (r.ptr) cudaHostUnregister(r.ptr);
    for (auto &pi : ctx->peer_images)
    {
        if (pi.owned) cudaFree(pi.ptr);
        else cudaIpcCloseMemHandle(pi.ptr);
    }
    if (ctx->abort_stream) { cudaStreamSynchronize(ctx->abort_stream); cudaStreamDestroy(ctx->abort_stream); }
    if (ctx->abort_host) cudaFreeHost(ctx->abort_host);
    if (ctx->stage_out.p) cudaFreeHost(ctx->stage_out.p);
    if (ctx->stage_in.p) cudaFreeHost(ctx->stage_in.p);
    if (ctx->stage_rad.p) cudaFreeHost(ctx->stage_rad.p);
    if (ctx->ev_stage) cudaEventDestroy(ctx->ev_stage);
    if (ctx->abort_dev) cudaFree(ctx->abort_dev);
    DevBuf *bufs[] = {&ctx->tile_centres, &ctx->scene_info, &ctx->aos, &ctx->rec, &ctx->cullrec, &ctx->ccounts, &ctx->coffsets, &ctx->cidx, &ctx->bin_count, &ctx->bin_off,
                      &ctx->bin_idx, &ctx->slice_dev, &ctx->hist, &ctx->queue, &ctx->stats, &ctx->counter, &ctx->rowcost, &ctx->scan_tmp, &ctx->cell_slot, &ctx->partial, &ctx->cwide,
                      &ctx->out_image, &ctx->out_rad, &ctx->tile_aos, &ctx->tile_off, &ctx->lit_offsets, &ctx->lit_counts, &ctx->lit_idx};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
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

int vrt_cuda_set_slice(vrt_cuda_ctx *ctx, int slice)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (slice != 0 && slice != 8 && slice != 16 && slice != 32 && slice != 64 && slice != 128 && slice != 256)
        return fail(ctx, VRT_CUDA_E_INVALID, "slice must be 0 (automatic), 8, 16, 32, 64, 128 or 256");
    ctx->tune_slice = slice;
    return 0;
}

int vrt_cuda_auto_slice(vrt_cuda_ctx *ctx, double share, int *slice_out)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (!ctx->have_lists) return fail(ctx, VRT_CUDA_E_STATE, "no lists: call vrt_cuda_tile first");
    if (!slice_out || !(share > 0.0 && share <= 1.0)) return fail(ctx, VRT_CUDA_E_INVALID, "share must be in (0, 1]");
    *slice_out = auto_slice(ctx, share);
    return 0;
}

int vrt_cuda_sync(vrt_cuda_ctx *ctx)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Page-locked staging owned by the context (grow-only).  Failure is not an error: the caller falls back to the plain copy.
static bool reserve_stage(vrt_cuda_ctx::HostStage &st, size_t bytes)
{
    if (bytes <= st.cap) return true;
    if (st.p) cudaFreeHost(st.p);
    st = vrt_cuda_ctx::HostStage{};
    const size_t want = bytes + bytes / 8 + 4096;
    void *p = nullptr, *d = nullptr;
    if (cudaHostAlloc(&p, want, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cudaHostGetDevicePointer(&d, p, 0) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(p); return false; }
    st.p = p;
    st.dev = d;
    st.cap = want;
    return true;
}

// memcpy by a few threads (one per MB, at most 12): a single core moves ~10 GB/s, the staging copies want PCIe rate
static void parallel_copy(void *dst, const void *src, size_t bytes)
{
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nt = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(12, hw), bytes >> 20));
    if (nt == 1) { std::memcpy(dst, src, bytes); return; }
    const size_t per = ((bytes + nt - 1) / nt + 63) & ~(size_t)63;
    std::vector<std::thread> th;
    size_t done_to = std::min(bytes, per); // [0, per) is this thread's share; [per, started_to) belongs to the helpers that started
    size_t started_to = done_to;
    try
    {
        th.reserve(nt - 1);
        for (size_t t = 1; t < nt; ++t)
        {
            const size_t off = std::min(bytes, t * per), len = std::min(bytes, off + per) - off;
            if (len) th.emplace_back([=] { std::memcpy((char *)dst + off, (const char *)src + off, len); });
            started_to = off + len;
        }
    }
    catch (...)
    {
        // (no thread to be had: nothing may escape through the C ABI -- this thread copies what no helper took)
    }
    std::memcpy(dst, src, done_to);
    if (started_to < bytes) std::memcpy((char *)dst + started_to, (const char *)src + started_to, bytes - started_to);
    for (auto &t : th) t.join();
}

// What kind of host memory is [p, p + bytes)?  Page-locked as a whole (CUDA copies from it asynchronously, kernels can be handed
// its device alias), pageable, or MIXED: only part of it lies inside a registration -- two small heap buffers that share a page,
// one of them registered, are enough -- which cudaMemcpy rejects ("invalid argument") and a kernel must not be pointed at.
enum class HostKind { pageable, locked, mixed };
static HostKind host_kind(const void *p, size_t bytes)
{
    if (!bytes) return HostKind::pageable;
    const char *probe[3] = {(const char *)p, (const char *)p + bytes / 2, (const char *)p + bytes - 1};
    int locked = 0;
    for (const char *q : probe)
    {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, q) == cudaSuccess && at.type == cudaMemoryTypeHost) ++locked;
        cudaGetLastError();
    }
    return locked == 3 ? HostKind::locked : (locked == 0 ? HostKind::pageable : HostKind::mixed);
}

// Page-lock a caller buffer once per (pointer, size); the registration is kept until the slot is reused or the context is
// destroyed.  Failure is not an error: the copy then takes the driver's staged path.
static void pin_host(vrt_cuda_ctx *ctx, const void *p, size_t bytes)
{
    if (!ctx->pin_mode || !p || bytes < (1u << 20)) return;
    for (auto &r : ctx->pinned)
        if (r.ptr == p && r.bytes >= bytes) return;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) return; // already pinned / managed by the caller
    cudaGetLastError();
    auto &slot = ctx->pinned[ctx->pinned_next];
    if (slot.ptr)
    {
        cudaStreamSynchronize(ctx->stream);
        cudaHostUnregister(slot.ptr);
        slot = vrt_cuda_ctx::Pinned{};
    }
    // registration covers whole pages
    const uintptr_t lo = (uintptr_t)p & ~(uintptr_t)4095, hi = ((uintptr_t)p + bytes + 4095) & ~(uintptr_t)4095;
    if (cudaHostRegister((void *)lo, hi - lo, cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess)
    {
        slot.ptr = (void *)lo;
        slot.bytes = hi - lo;
        ctx->pinned_next = (ctx->pinned_next + 1) % 4;
    }
    else cudaGetLastError();
}

int vrt_cuda_set_host_pinning(vrt_cuda_ctx *ctx, int on)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    ctx->pin_mode = on ? 1 : 0;
    if (!on)
    {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        for (auto &r : ctx->pinned)
        {
            if (r.ptr) cudaHostUnregister(r.ptr);
            r = vrt_cuda_ctx::Pinned{};
        }
    }
    return 0;
}

int vrt_cuda_pin_buffer(vrt_cuda_ctx *ctx, void *p, uint64_t bytes)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (!p || !bytes) return fail(ctx, VRT_CUDA_E_INVALID, "nothing to pin");
    CU(cudaSetDevice(ctx->device));
    const uintptr_t lo = (uintptr_t)p & ~(uintptr_t)4095, hi = ((uintptr_t)p + bytes + 4095) & ~(uintptr_t)4095;
    CU(cudaHostRegister((void *)lo, hi - lo, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return 0;
}

int vrt_cuda_unpin_buffer(vrt_cuda_ctx *ctx, void *p)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaHostUnregister((void *)((uintptr_t)p & ~(uintptr_t)4095)));
    return 0;
}

// ---- peer images: a frame buffer on one GPU that the other ranks' render kernels store into (CUDA IPC over NVLink) ----
static_assert(sizeof(cudaIpcMemHandle_t) == VRT_CUDA_PEER_HANDLE_BYTES, "handle size");

int vrt_cuda_peer_image_create(vrt_cuda_ctx *ctx, uint64_t bytes, void **image_dev_out, unsigned char *handle_out)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (!bytes || !image_dev_out || !handle_out) return fail(ctx, VRT_CUDA_E_INVALID, "peer image: size, pointer and handle are required");
    CU(cudaSetDevice(ctx->device));
    void *p = nullptr;
    // an allocation of its own (not a slice of a pool): the handle names a whole cudaMalloc block
    if (cudaMalloc(&p, bytes) != cudaSuccess)
    {
        cudaGetLastError();
        return fail(ctx, VRT_CUDA_E_NOMEM, "peer image: cannot allocate %llu bytes", (unsigned long long)bytes);
    }
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        cudaFree(p);
        return fail(ctx, VRT_CUDA_E_CUDA, "peer image: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    std::memcpy(handle_out, &h, sizeof(h));
    ctx->peer_images.push_back({p, true});
    *image_dev_out = p;
    return 0;
}

int vrt_cuda_peer_image_open(vrt_cuda_ctx *ctx, const unsigned char *handle, void **image_dev_out)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (!handle || !image_dev_out) return fail(ctx, VRT_CUDA_E_INVALID, "peer image: handle and pointer are required");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        return fail(ctx, VRT_CUDA_E_CUDA, "peer image: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    }
    ctx->peer_images.push_back({p, false});
    *image_dev_out = p;
    return 0;
}

int vrt_cuda_peer_image_close(vrt_cuda_ctx *ctx, void *image_dev)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    for (size_t i = 0; i < ctx->peer_images.size(); ++i)
        if (ctx->peer_images[i].ptr == image_dev)
        {
            CU(cudaSetDevice(ctx->device));
            CU(cudaStreamSynchronize(ctx->stream));
            const bool owned = ctx->peer_images[i].owned;
            ctx->peer_images.erase(ctx->peer_images.begin() + (long)i);
            if (owned) CU(cudaFree(image_dev));
            else CU(cudaIpcCloseMemHandle(image_dev));
            return 0;
        }
    return fail(ctx, VRT_CUDA_E_INVALID, "peer image: not an image of this context");
}

int vrt_cuda_abort(vrt_cuda_ctx *ctx, int on)
{
    if (!ctx || !ctx->abort_host) return VRT_CUDA_E_INVALID;
    // (no `CU`: this may run on a thread other than the one driving the context, and must not touch ctx->err)
    if (cudaSetDevice(ctx->device) != cudaSuccess) return VRT_CUDA_E_CUDA;
    *ctx->abort_host = on ? 1u : 0u;
    if (cudaMemcpyAsync(ctx->abort_dev, ctx->abort_host, sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->abort_stream) != cudaSuccess) return VRT_CUDA_E_CUDA;
    if (cudaStreamSynchronize(ctx->abort_stream) != cudaSuccess) return VRT_CUDA_E_CUDA;
    return 0;
}

static int set_gaussians_impl(vrt_cuda_ctx *ctx, const float *aos, uint64_t n, cudaMemcpyKind kind)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (n && !aos) return fail(ctx, VRT_CUDA_E_INVALID, "aos is NULL");
    if (n > 0x7FFFFFF0ull) return fail(ctx, VRT_CUDA_E_INVALID, "too many Gaussians (device indices are 32-bit, cull records are addressed as 2 * index)");
    CU(cudaSetDevice(ctx->device));
    if (int rc = reserve(ctx, ctx->aos, std::max<size_t>(n, 1) * 40)) return rc;
    if (n && kind == cudaMemcpyHostToDevice) pin_host(ctx, aos, n * 40);
    constexpr size_t STAGE_MIN = 256u << 10, STAGE_CHUNK = 8u << 20;
    const HostKind hk = (n && kind == cudaMemcpyHostToDevice) ? host_kind(aos, n * 40) : HostKind::locked;
    if (hk == HostKind::mixed && !(ctx->ev_stage && reserve_stage(ctx->stage_in, n * 40)))
        return fail(ctx, VRT_CUDA_E_NOMEM, "the scene buffer is only partly page-locked and no staging buffer could be allocated");
    if (n && kind == cudaMemcpyHostToDevice && (hk == HostKind::mixed || (n * 40 >= STAGE_MIN && hk == HostKind::pageable)) && ctx->ev_stage &&
        reserve_stage(ctx->stage_in, n * 40))
    {
        // pageable scene: host threads copy it into the context's page-locked staging chunk by chunk, every chunk's upload
        // overlaps the next chunk's host copy; `aos` is not referenced after this returns
        CU(cudaEventSynchronize(ctx->ev_stage)); // (the previous upload out of the staging)
        const size_t bytes = n * 40;
        for (size_t off = 0; off < bytes; off += STAGE_CHUNK)
        {
            const size_t len = std::min(STAGE_CHUNK, bytes - off);
            parallel_copy((char *)ctx->stage_in.p + off, (const char *)aos + off, len);
            CU(cudaMemcpyAsync((char *)ctx->aos.p + off, (const char *)ctx->stage_in.p + off, len, cudaMemcpyHostToDevice, ctx->stream));
        }
        CU(cudaEventRecord(ctx->ev_stage, ctx->stream));
    }
    else if (n) CU(cudaMemcpyAsync(ctx->aos.p, aos, n * 40, kind, ctx->stream));
    // a copy out of page-locked memory is truly asynchronous: the caller may reuse `aos` as soon as this returns
    if (n && kind == cudaMemcpyHostToDevice && ctx->pin_mode) CU(cudaStreamSynchronize(ctx->stream));
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

int vrt_cuda_set_band_tuning(vrt_cuda_ctx *ctx, int ctas_per_sm)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (ctas_per_sm != 4 && ctas_per_sm != 5) return fail(ctx, VRT_CUDA_E_INVALID, "the banded kernel runs 4 or 5 CTAs per SM");
    ctx->tune_band_ctas = ctas_per_sm;
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

int vrt_cuda_approx_rate(vrt_cuda_ctx *ctx, int fn, double *values_per_s_out)
{
    if (!ctx || !values_per_s_out) return VRT_CUDA_E_INVALID;
    if (fn < 0 || fn > VRT_CUDA_FN_SPLINE_EXP) return fail(ctx, VRT_CUDA_E_INVALID, "unknown function id %d", fn);
    CU(cudaSetDevice(ctx->device));
    if (int rc = reserve(ctx, ctx->counter, sizeof(uint32_t) * 4)) return rc;
    const int iters = 1024, blocks = ctx->sm_count * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep)
    {
        CU(cudaEventRecord(ctx->ev[2], ctx->stream));
        switch (fn)
        {
        case VRT_CUDA_FN_SPLINE_ERF: launch_approx_rate<VRT_CUDA_FN_SPLINE_ERF>(ctx, blocks, iters); break;
        case VRT_CUDA_FN_SPLINE_ERF_MIRROR: launch_approx_rate<VRT_CUDA_FN_SPLINE_ERF_MIRROR>(ctx, blocks, iters); break;
        case VRT_CUDA_FN_TAYLOR_ERF: launch_approx_rate<VRT_CUDA_FN_TAYLOR_ERF>(ctx, blocks, iters); break;
        case VRT_CUDA_FN_AS_ERF: launch_approx_rate<VRT_CUDA_FN_AS_ERF>(ctx, blocks, iters); break;
        case VRT_CUDA_FN_ERF: launch_approx_rate<VRT_CUDA_FN_ERF>(ctx, blocks, iters); break;
        case VRT_CUDA_FN_EXP: launch_approx_rate<VRT_CUDA_FN_EXP>(ctx, blocks, iters); break;
        case VRT_CUDA_FN_FAST_EXP: launch_approx_rate<VRT_CUDA_FN_FAST_EXP>(ctx, blocks, iters); break;
        default: launch_approx_rate<VRT_CUDA_FN_SPLINE_EXP>(ctx, blocks, iters); break;
        }
        CU(cudaGetLastError());
        CU(cudaEventRecord(ctx->ev[3], ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
        if (rep > 0 && ms < best) best = ms;
    }
    *values_per_s_out = (double)blocks * 256.0 * 8.0 * iters / (best * 1e-3);
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

// host-side constants of the binning grid (k1_bin.cuh): the two families of planes through the apex, one per screen axis
static BinGrid make_bin_grid(const FrameGeom &G)
{
    BinGrid B{};
    B.nbx = (G.ncx + BIN_CX - 1) / BIN_CX;
    B.nby = (G.ncy + BIN_CY - 1) / BIN_CY;
    float Wv[3];
    for (int i = 0; i < 3; ++i) Wv[i] = G.inv3[i] - G.origin[i];
    auto cross = [](const float *a, const float *b, float *r) {
        r[0] = a[1] * b[2] - a[2] * b[1];
        r[1] = a[2] * b[0] - a[0] * b[2];
        r[2] = a[0] * b[1] - a[1] * b[0];
    };
    auto dot = [](const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    cross(G.inv1, G.inv0, B.px); // x planes contain U = inv1 and the ray u R + Wv: n(u) = U x (u R + Wv)
    cross(G.inv1, Wv, B.qx);
    cross(G.inv0, G.inv1, B.py); // y planes contain R = inv0: n(v) = R x (v U + Wv)
    cross(G.inv0, Wv, B.qy);
    B.pp_x = dot(B.px, B.px); B.pq_x = dot(B.px, B.qx); B.qq_x = dot(B.qx, B.qx);
    B.pp_y = dot(B.py, B.py); B.pq_y = dot(B.py, B.qy); B.qq_y = dot(B.qy, B.qy);
    B.u_min = -1.f; B.u_max = -1.f + (float)(G.W - 1) / G.half_w;
    B.v_min = -1.f; B.v_max = -1.f + (float)(G.H - 1) / G.half_h;
    return B;
}

// K1 for per-cell lists: binning + the fused cull / sort pass.  src_tiles = true: the candidates of a cell are the records of
// its reference tile (caller-supplied tiles_t), not a bin list.  Capacities come from the previous frame (or a first guess) and
// are checked by the caller after the frame's one synchronisation.
static int enqueue_cell_lists(vrt_cuda_ctx *ctx, const FrameGeom &G, uint64_t n_records, bool src_tiles, const uint32_t *tile_off)
{
    const uint64_t cells = (uint64_t)G.ncx * G.ncy;
    if (cells > 0x7FFFFFFFull / 32) return fail(ctx, VRT_CUDA_E_INVALID, "image too large");
    TileStats *stats = (TileStats *)ctx->stats.p;
    LeafArgs L{};
    L.cullrec = (const float4 *)ctx->cullrec.p;
    if (!src_tiles)
    {
        const BinGrid B = make_bin_grid(G);
        const uint64_t nb = (uint64_t)B.nbx * B.nby;
        if (int rc = reserve(ctx, ctx->bin_count, sizeof(uint32_t) * nb)) return rc;
        if (int rc = reserve(ctx, ctx->bin_off, sizeof(uint32_t) * (nb + 1))) return rc;
        const uint64_t want = std::max<uint64_t>(ctx->bin_hint, std::max<uint64_t>(4 * n_records, 1u << 20));
        if (int rc = reserve(ctx, ctx->bin_idx, sizeof(uint32_t) * want)) return rc;
        const uint64_t bin_cap = std::min<uint64_t>(ctx->bin_idx.cap / sizeof(uint32_t), 0xFFFFFFF0ull);
        const unsigned grid = (unsigned)((n_records + 255) / 256);
        CU(cudaMemsetAsync(ctx->bin_count.p, 0, sizeof(uint32_t) * nb, ctx->stream));
        if (n_records) k1_bin<false><<<grid, 256, 0, ctx->stream>>>(G, B, (const float4 *)ctx->cullrec.p, (uint32_t)n_records, (uint32_t *)ctx->bin_count.p, nullptr, nullptr, 0, &stats->bin_entries);
        ctx->launches++;
        if (int rc = scan_u32(ctx, (const uint32_t *)ctx->bin_count.p, (uint32_t *)ctx->bin_off.p, (uint32_t)nb)) return rc;
        CU(cudaMemsetAsync(ctx->bin_count.p, 0, sizeof(uint32_t) * nb, ctx->stream));
        if (n_records) k1_bin<true><<<grid, 256, 0, ctx->stream>>>(G, B, (const float4 *)ctx->cullrec.p, (uint32_t)n_records, (uint32_t *)ctx->bin_count.p, (const uint32_t *)ctx->bin_off.p,
                                                   (uint32_t *)ctx->bin_idx.p, bin_cap, nullptr);
        ctx->launches++;
        L.bin_off = (const uint32_t *)ctx->bin_off.p;
        L.bin_count = (const uint32_t *)ctx->bin_count.p;
        L.bin_idx = (const uint32_t *)ctx->bin_idx.p;
        L.bin_cap = bin_cap;
        L.nbx = B.nbx;
    }
    else L.tile_off = tile_off;
    if (int rc = reserve(ctx, ctx->ccounts, sizeof(uint32_t) * cells)) return rc;
    if (int rc = reserve(ctx, ctx->coffsets, sizeof(uint32_t) * (cells + 1))) return rc;
    // first guess for the index array: 96 entries per cell of the band (BASELINE configs 4/5 need ~70-85)
    const uint64_t band_cells = (uint64_t)G.ncx * ((uint64_t)(G.row_end - G.row_begin + CELL_H - 1) / CELL_H + 1);
    const uint64_t want = std::max<uint64_t>(ctx->idx_hint, band_cells * 96 + (1u << 16));
    if (int rc = reserve(ctx, ctx->cidx, sizeof(uint32_t) * want)) return rc;
    L.list_off = (uint32_t *)ctx->coffsets.p;
    L.list_cnt = (uint32_t *)ctx->ccounts.p;
    L.list_idx = (uint32_t *)ctx->cidx.p;
    L.cursor = &stats->leaf_entries;
    L.idx_cap = std::min<uint64_t>(ctx->cidx.cap / sizeof(uint32_t), 0xFFFFFFF0ull);
    L.n_cells = (uint32_t)cells;
    if (int rc = reserve(ctx, ctx->cwide, (size_t)cells)) return rc;
    L.list_wide = (uint8_t *)ctx->cwide.p;
    L.wide_frac = ctx->long_wide;
    // Only the cell rows that intersect the rendered row band are launched (a rank of an N-GPU frame pays 1/N of the leaf pass,
    // not the launch of 65 536 CTAs that return at once); every other cell's list is empty by the memset.
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
    const bool whole = cyb == 0 && cye == G.ncy;
    if (!whole)
    {
        CU(cudaMemsetAsync(ctx->ccounts.p, 0, sizeof(uint32_t) * cells, ctx->stream));
        CU(cudaMemsetAsync(ctx->coffsets.p, 0, sizeof(uint32_t) * cells, ctx->stream));
    }
    if (cye > cyb)
    {
        L.cell_begin = (uint32_t)cyb * (uint32_t)G.ncx;
        L.n_cells = (uint32_t)cye * (uint32_t)G.ncx;
        const unsigned grid = (unsigned)((L.n_cells - L.cell_begin + LEAF_WARPS - 1) / LEAF_WARPS);
        if (src_tiles) k1_leaf<1><<<grid, LEAF_WARPS * 32, 0, ctx->stream>>>(G, L);
        else k1_leaf<0><<<grid, LEAF_WARPS * 32, 0, ctx->stream>>>(G, L);
    }
    ctx->launches++;
    ctx->n_lists = (uint32_t)cells;
    ctx->lists_sorted = true;
    CU(cudaGetLastError());
    return 0;
}

// Most list entries a frame may hold: device offsets are 32-bit.  VRT_CUDA_MAX_LIST_ENTRIES lowers the limit so that the
// tests can reach the refusal without building four billion entries.
static uint64_t max_list_entries()
{
    static const uint64_t limit = [] {
        const char *e = std::getenv("VRT_CUDA_MAX_LIST_ENTRIES");
        const unsigned long long v = e ? std::strtoull(e, nullptr, 10) : 0ull;
        return (uint64_t)((v > 0 && v < 0xFFFFFFF0ull) ? v : 0xFFFFFFF0ull);
    }();
    return limit;
}

// after the frame's synchronisation: did the bin lists and the index array fit?  (false: capacities were raised, run again)
static int cell_lists_fit(vrt_cuda_ctx *ctx, const TileStats &ts, bool &fit)
{
    fit = true;
    if (ts.leaf_entries > max_list_entries() || ts.bin_entries > max_list_entries())
        return fail(ctx, VRT_CUDA_E_NOMEM, "too many list entries in this frame (%llu cell entries, %llu bin entries; device offsets are 32-bit): render it in row bands or tighten bound_sigmas",
                    (unsigned long long)ts.leaf_entries, (unsigned long long)ts.bin_entries);
    if (ts.bin_entries > ctx->bin_idx.cap / sizeof(uint32_t))
    {
        ctx->bin_hint = ts.bin_entries + ts.bin_entries / 4 + 4096;
        fit = false;
    }
    else ctx->bin_hint = std::max<uint64_t>(ctx->bin_hint / 2, ts.bin_entries + ts.bin_entries / 8 + 4096);
    if (ts.leaf_entries > ctx->cidx.cap / sizeof(uint32_t))
    {
        ctx->idx_hint = ts.leaf_entries + ts.leaf_entries / 4 + 4096;
        fit = false;
    }
    else ctx->idx_hint = std::max<uint64_t>(ctx->idx_hint / 2, ts.leaf_entries + ts.leaf_entries / 8 + 4096);
    ctx->n_entries = ts.leaf_entries;
    return 0;
}

// K0 + K1 for one frame: records, lists of the frame's list mode (depth-sorted for the per-cell modes), work queue
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
    ctx->launches = 0;
    const uint64_t N = ctx->n_gauss;
    if (int rc = upload_geom(ctx, G, cxs, cys)) return rc;
    ctx->geom = G;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));

    if (int rc = reserve(ctx, ctx->rec, sizeof(Rec) * std::max<uint64_t>(N, 1))) return rc;
    if (int rc = reserve(ctx, ctx->cullrec, 2 * sizeof(float4) * std::max<uint64_t>(N, 1))) return rc;
    if (int rc = reserve(ctx, ctx->scene_info, sizeof(uint32_t) * 4)) return rc;
    if (int rc = reserve(ctx, ctx->stats, sizeof(TileStats))) return rc;
    CU(cudaMemsetAsync(ctx->scene_info.p, 0, sizeof(uint32_t) * 4, ctx->stream));
    if (N)
    {
        k0_prepare<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(G, (const float *)ctx->aos.p, N, (Rec *)ctx->rec.p, (float4 *)ctx->cullrec.p,
                                                                         (uint32_t *)ctx->scene_info.p);
        ctx->launches++;
    }
    ctx->tiled_list_mode = frame->flags & VRT_CUDA_LIST_MASK;
    ctx->tiled_tiles_x = G.tiles_x;
    ctx->tiled_tiles_y = G.tiles_y;
    ctx->tiled_bound = G.bound_k;
    ctx->lists_sorted = false;
    TileStats ts{};

    if (G.list_kind == 0)
    {
        // bounded per-cell lists: no host round trip inside; capacities from the previous frame, checked afterwards
        for (int attempt = 0;; ++attempt)
        {
            CU(cudaMemsetAsync(ctx->stats.p, 0, sizeof(TileStats), ctx->stream));
            if (int rc = enqueue_cell_lists(ctx, G, N, false, nullptr)) return rc;
            const uint64_t queue_cap = (uint64_t)G.ncx * G.ncy + ctx->cidx.cap / sizeof(uint32_t) / SLICE_MIN + 1;
            if (int rc = build_queue(ctx, std::min<uint64_t>(queue_cap, 0xFFFFFFF0ull))) return rc;
            CU(cudaEventRecord(ctx->ev[1], ctx->stream));
            if (int rc = finish_tile(ctx, ts)) return rc;
            bool fit = true;
            if (int rc = cell_lists_fit(ctx, ts, fit)) return rc;
            if (fit) break;
            if (attempt >= 2) return fail(ctx, VRT_CUDA_E_NOMEM, "list capacities did not converge");
        }
    }
    else
    {
        CU(cudaMemsetAsync(ctx->stats.p, 0, sizeof(TileStats), ctx->stream));
        if (G.list_kind == 2)
        {
            // single list: all Gaussians, contiguous
            if (int rc = reserve(ctx, ctx->coffsets, sizeof(uint32_t) * 2)) return rc;
            if (int rc = reserve(ctx, ctx->ccounts, sizeof(uint32_t) * 2)) return rc;
            const uint32_t off[2] = {0u, (uint32_t)N}, cnt[2] = {(uint32_t)N, 0u};
            CU(cudaMemcpyAsync(ctx->coffsets.p, off, sizeof(off), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemcpyAsync(ctx->ccounts.p, cnt, sizeof(cnt), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream)); // off[] / cnt[] are on the stack
            ctx->n_lists = 1;
            ctx->n_entries = N;
        }
        else
        {
            // the literal membership of tile_gaussians, one list per reference tile, ascending Gaussian index (count, scan, write)
            const uint32_t nt = (uint32_t)(G.tiles_x * G.tiles_y);
            if (int rc = reserve(ctx, ctx->ccounts, sizeof(uint32_t) * nt)) return rc;
            if (int rc = reserve(ctx, ctx->coffsets, sizeof(uint32_t) * (nt + 1))) return rc;
            if (int rc = reserve(ctx, ctx->scan_tmp, sizeof(unsigned long long))) return rc;
            k1_cull_tiles<false><<<nt, 256, 0, ctx->stream>>>(G, (const Rec *)ctx->rec.p, (const float4 *)ctx->cullrec.p, (uint32_t)N, (uint32_t *)ctx->ccounts.p, nullptr, nullptr, nt);
            // the grand total in 64 bits: 32-bit offsets wrap silently past 2^32 entries (tiles x N can get there)
            CU(cudaMemsetAsync(&((TileStats *)ctx->stats.p)->leaf_entries, 0, sizeof(unsigned long long), ctx->stream));
            k1_total64<<<(nt + 255) / 256, 256, 0, ctx->stream>>>((const uint32_t *)ctx->ccounts.p, nt, &((TileStats *)ctx->stats.p)->leaf_entries);
            k1_scan<<<1, 1024, 0, ctx->stream>>>((const uint32_t *)ctx->ccounts.p, (uint32_t *)ctx->coffsets.p, nt);
            unsigned long long total = 0;
            CU(cudaMemcpyAsync(&total, &((TileStats *)ctx->stats.p)->leaf_entries, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            if (total > max_list_entries())
                return fail(ctx, VRT_CUDA_E_NOMEM, "the literal tile lists hold %llu entries (device offsets are 32-bit): use a *_BOUND list mode or fewer tiles", total);
            if (int rc = reserve(ctx, ctx->cidx, sizeof(uint32_t) * std::max<uint64_t>(total, 1))) return rc;
            k1_cull_tiles<true><<<nt, 256, 0, ctx->stream>>>(G, (const Rec *)ctx->rec.p, (const float4 *)ctx->cullrec.p, (uint32_t)N, nullptr, (const uint32_t *)ctx->coffsets.p, (uint32_t *)ctx->cidx.p, nt);
            ctx->launches += 4;
            ctx->n_lists = nt;
            ctx->n_entries = total;
        }
        CU(cudaGetLastError());
        if (int rc = build_queue(ctx, 0)) return rc;
        CU(cudaEventRecord(ctx->ev[1], ctx->stream));
        if (int rc = finish_tile(ctx, ts)) return rc;
    }
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
    std::swap(ctx->ccounts, ctx->lit_counts);
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
    ctx->tiled_list_mode = lm; // what the caller asked for (the visible subsets are an implementation detail)
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
    if (total > 0x7FFFFFF0ull) return fail(ctx, VRT_CUDA_E_INVALID, "tile lists too long (device indices are 32-bit)");
    if (total && !aos_concat) return fail(ctx, VRT_CUDA_E_INVALID, "aos_concat is NULL");
    std::vector<uint32_t> off32(n_tiles + 1), cnt32(n_tiles + 1, 0u);
    for (uint64_t t = 0; t <= n_tiles; ++t)
    {
        if (t && offsets[t] < offsets[t - 1]) return fail(ctx, VRT_CUDA_E_INVALID, "offsets must be non-decreasing");
        off32[t] = (uint32_t)offsets[t];
        if (t) cnt32[t - 1] = (uint32_t)(offsets[t] - offsets[t - 1]);
    }
    ctx->have_lists = false;
    ctx->literal = false;
    ctx->launches = 0;
    if (int rc = upload_geom(ctx, G, cxs, cys)) return rc;
    ctx->geom = G;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (int rc = reserve(ctx, ctx->scene_info, sizeof(uint32_t) * 4)) return rc;
    CU(cudaMemsetAsync(ctx->scene_info.p, 0, sizeof(uint32_t) * 4, ctx->stream));
    if (int rc = reserve(ctx, ctx->tile_aos, std::max<uint64_t>(total, 1) * 40)) return rc;
    if (int rc = reserve(ctx, ctx->rec, sizeof(Rec) * std::max<uint64_t>(total, 1))) return rc;
    if (int rc = reserve(ctx, ctx->cullrec, 2 * sizeof(float4) * std::max<uint64_t>(total, 1))) return rc;
    if (int rc = reserve(ctx, ctx->coffsets, sizeof(uint32_t) * (n_tiles + 1))) return rc;
    if (int rc = reserve(ctx, ctx->ccounts, sizeof(uint32_t) * (n_tiles + 1))) return rc;
    if (total) CU(cudaMemcpyAsync(ctx->tile_aos.p, aos_concat, total * 40, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->coffsets.p, off32.data(), sizeof(uint32_t) * (n_tiles + 1), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->ccounts.p, cnt32.data(), sizeof(uint32_t) * (n_tiles + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (total)
    {
        k0_prepare<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(G, (const float *)ctx->tile_aos.p, total, (Rec *)ctx->rec.p, (float4 *)ctx->cullrec.p,
                                                                             (uint32_t *)ctx->scene_info.p);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    ctx->n_lists = (uint32_t)n_tiles;
    ctx->n_entries = total;
    ctx->lists_from_host = true;
    ctx->lists_sorted = false;
    if (int rc = reserve(ctx, ctx->stats, sizeof(TileStats))) return rc;
    CU(cudaMemsetAsync(ctx->stats.p, 0, sizeof(TileStats), ctx->stream));
    TileStats ts{};
    if (int rc = build_queue(ctx, 0)) return rc; // (synchronises: off32 / cnt32 have been consumed)
    if (total && !(frame->flags & VRT_CUDA_NO_SKIP))
    {
        // visible lists (see vrt_cuda_tile): every 8x4 cell scans the record range of its own tile once (k1_leaf)
        if (int rc = keep_literal(ctx)) return rc;
        FrameGeom V = G;
        V.list_kind = 0;
        V.use_ref = 0;
        V.use_bound = 1;
        V.bound_k = VISIBLE_SIGMAS;
        ctx->geom = V; // (same tile-centre buffer)
        for (int attempt = 0;; ++attempt)
        {
            CU(cudaMemsetAsync(ctx->stats.p, 0, sizeof(TileStats), ctx->stream));
            if (int rc = enqueue_cell_lists(ctx, V, total, true, (const uint32_t *)ctx->lit_offsets.p)) return rc;
            const uint64_t queue_cap = (uint64_t)V.ncx * V.ncy + ctx->cidx.cap / sizeof(uint32_t) / SLICE_MIN + 1;
            if (int rc = build_queue(ctx, std::min<uint64_t>(queue_cap, 0xFFFFFFF0ull))) return rc;
            CU(cudaEventRecord(ctx->ev[1], ctx->stream));
            if (int rc = finish_tile(ctx, ts)) return rc;
            bool fit = true;
            if (int rc = cell_lists_fit(ctx, ts, fit)) return rc;
            if (fit) break;
            if (attempt >= 2) return fail(ctx, VRT_CUDA_E_NOMEM, "list capacities did not converge");
        }
        ctx->literal = true;
    }
    else
    {
        CU(cudaEventRecord(ctx->ev[1], ctx->stream));
        if (int rc = finish_tile(ctx, ts)) return rc;
    }
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
    const DevBuf &counts = ctx->literal ? ctx->lit_counts : ctx->ccounts;
    const DevBuf &indices = ctx->literal ? ctx->lit_idx : ctx->cidx;
    if (n_cells_out) *n_cells_out = n_lists;
    if (n_entries_out) *n_entries_out = n_entries;
    std::vector<uint32_t> cnt(n_lists);
    if ((counts_out || idx_out) && n_lists)
    {
        CU(cudaMemcpyAsync(cnt.data(), counts.p, sizeof(uint32_t) * n_lists, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (counts_out)
    {
        if (counts_cap < n_lists) return fail(ctx, VRT_CUDA_E_INVALID, "counts_cap too small");
        for (uint32_t i = 0; i < n_lists; ++i) counts_out[i] = cnt[i];
    }
    if (idx_out)
    {
        if (idx_cap < n_entries) return fail(ctx, VRT_CUDA_E_INVALID, "idx_cap too small");
        if (kind == 2 || (ctx->lists_from_host && kind == 1))
        {
            for (uint64_t i = 0; i < n_entries; ++i) idx_out[i] = (uint32_t)i;
        }
        else if (n_entries)
        {
            // lists are (offset, count) pairs placed anywhere in the index array: return them concatenated in list order
            std::vector<uint32_t> off(n_lists), all(n_entries);
            CU(cudaMemcpyAsync(off.data(), offsets.p, sizeof(uint32_t) * n_lists, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaMemcpyAsync(all.data(), indices.p, sizeof(uint32_t) * n_entries, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            uint64_t at = 0;
            for (uint32_t i = 0; i < n_lists; ++i)
            {
                if ((uint64_t)off[i] + cnt[i] > n_entries || at + cnt[i] > n_entries) return fail(ctx, VRT_CUDA_E_STATE, "inconsistent list table");
                std::memcpy(idx_out + at, all.data() + off[i], sizeof(uint32_t) * cnt[i]);
                at += cnt[i];
            }
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

// statistics of the last tile + render (waits for the stream)
static int read_stats(vrt_cuda_ctx *ctx, vrt_cuda_stats *stats)
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
    stats->terms_terminated = (double)ts.terms_term;
    stats->ms_tile = ctx->ms_tile;
    CU(cudaEventElapsedTime(&stats->ms_render, ctx->ev[2], ctx->ev[3]));
    stats->ms_total = stats->ms_tile + stats->ms_render;
    stats->slice = (uint32_t)ctx->geom.slice;
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
    if (!ctx->lists_from_host)
    {
        // the lists belong to one list mode, tile count and bound: a frame that names another one would render with semantics
        // its flags do not describe
        const uint32_t lm = frame->flags & VRT_CUDA_LIST_MASK;
        const bool tiled = lm == VRT_CUDA_LIST_REFERENCE || lm == VRT_CUDA_LIST_REFERENCE_BOUND;
        const bool bound = lm == VRT_CUDA_LIST_REFERENCE_BOUND || lm == VRT_CUDA_LIST_BOUND;
        if (lm != ctx->tiled_list_mode) return fail(ctx, VRT_CUDA_E_STATE, "list mode differs from the tiled frame");
        if (tiled && ((int)frame->tiles_x != ctx->tiled_tiles_x || (int)frame->tiles_y != ctx->tiled_tiles_y))
            return fail(ctx, VRT_CUDA_E_STATE, "tile count differs from the tiled frame");
        if (bound && (frame->bound_sigmas > 0.f ? frame->bound_sigmas : 6.0f) != ctx->tiled_bound) return fail(ctx, VRT_CUDA_E_STATE, "bound_sigmas differs from the tiled frame");
    }
    const uint32_t approx_erf = frame->flags & VRT_CUDA_APPROX_ERF_MASK, approx_exp = frame->flags & VRT_CUDA_APPROX_EXP_MASK;
    if ((approx_erf || approx_exp) && (frame->flags & VRT_CUDA_DEPTH_WINDOW))
        return fail(ctx, VRT_CUDA_E_INVALID, "the depth window needs a saturating odd erf: not available with VRT_CUDA_APPROX_*");
    if (approx_exp > VRT_CUDA_APPROX_EXP_SPLINE) return fail(ctx, VRT_CUDA_E_INVALID, "unknown exp approximation");
    RenderArgs a{};
    a.geom = G;
    a.rec = (const Rec *)ctx->rec.p;
    a.list_off = (const uint32_t *)ctx->coffsets.p;
    a.list_cnt = (const uint32_t *)ctx->ccounts.p;
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
    a.terms_term = &((TileStats *)ctx->stats.p)->terms_term;
    a.scene_info = (const uint32_t *)ctx->scene_info.p;
    a.abort_flag = ctx->abort_dev;
    a.terminate = (frame->flags & VRT_CUDA_NO_TERMINATE) ? 0u : 1u;
    // banded evaluation is the default wherever the lists are depth-sorted index lists (every bounded mode and the visible
    // lists of the literal modes); VRT_CUDA_EVAL_ALL asks for every listed term, the approximation variants need it
    const bool explicit_window = (frame->flags & VRT_CUDA_DEPTH_WINDOW) != 0;
    if (explicit_window && (a.list_idx == nullptr || !ctx->lists_sorted))
        return fail(ctx, VRT_CUDA_E_INVALID, "VRT_CUDA_DEPTH_WINDOW needs depth-sorted per-cell lists: not available on literally walked lists (VRT_CUDA_NO_SKIP)");
    a.window = (a.list_idx != nullptr && ctx->lists_sorted && !(frame->flags & (VRT_CUDA_EVAL_ALL | VRT_CUDA_NO_SKIP)) && !approx_erf && !approx_exp) ? 1u : 0u;
    a.cell_slot = (const uint32_t *)ctx->cell_slot.p;
    if (ctx->n_split)
        if (int rc = reserve(ctx, ctx->partial, (size_t)ctx->n_split * 32 * sizeof(float4))) return rc;
    a.partial = (float4 *)ctx->partial.p;
    const bool bounded = G.use_bound != 0;
    a.skip_thresh = (frame->flags & VRT_CUDA_NO_SKIP) ? -1.f : ((bounded && !ctx->literal) ? std::exp2(-0.5f * G.bound_k * G.bound_k * LOG2E) : 0.f);
    a.image_vec16 = (image_dev && ((uintptr_t)image_dev & 15u) == 0 && (G.W & 3) == 0) ? 1u : 0u;
    a.quant_nearest = (frame->flags & VRT_CUDA_QUANT_NEAREST) ? 1u : 0u;
    a.alpha_from_w = (frame->flags & VRT_CUDA_ALPHA_FROM_W) ? 1u : 0u;
    CU(cudaMemsetAsync(ctx->counter.p, 0, sizeof(uint32_t) * 8, ctx->stream)); // [0] k2_render head, [1] k2_band, [3] k2_band_long
    CU(cudaMemsetAsync(&((TileStats *)ctx->stats.p)->terms_exec, 0, 2 * sizeof(unsigned long long), ctx->stream));
    CU(cudaMemsetAsync(&((TileStats *)ctx->stats.p)->terms_term, 0, sizeof(unsigned long long), ctx->stream));
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
        k3_combine<<<(unsigned)(((uint64_t)ncells * 32 + 255) / 256), 256, 0, ctx->stream>>>(a, ctx->cy_begin, ctx->cy_end);
        ctx->render_launches++;
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    if (stats) return read_stats(ctx, stats);
    return 0;
}

// running: optional pointer to the caller's `running` flag (the `const bool &running` of the reference's entries; one byte,
// written by another thread).  While the frame is in flight the host polls it; when it goes false the mapped abort word is
// raised, the persistent warps stop taking work items, and the call returns VRT_CUDA_INTERRUPTED with a partial image.
static int render_host(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame, uint32_t *image, float *radiance, vrt_cuda_stats *stats, const volatile unsigned char *running)
{
    if (!ctx) return VRT_CUDA_E_INVALID;
    if (!frame) return fail(ctx, VRT_CUDA_E_INVALID, "frame is NULL");
    CU(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)frame->width * frame->height;
    // A page-locked, mapped caller buffer (vrt_cuda_set_host_pinning / vrt_cuda_pin_buffer) is written by the render kernel
    // itself: K3's 16-byte stores go straight over PCIe while the frame is still being computed, and no copy-back remains.
    // Anything else is rendered into a device buffer and copied (staged by the driver when the memory is pageable).
    uint32_t *image_target = nullptr;
    float *rad_target = nullptr;
    bool image_staged = false, rad_mixed = false;
    if (image)
    {
        pin_host(ctx, image, npix * sizeof(uint32_t));
        const HostKind hk = host_kind(image, npix * sizeof(uint32_t));
        void *d = nullptr;
        if (hk == HostKind::locked && cudaHostGetDevicePointer(&d, image, 0) == cudaSuccess) image_target = (uint32_t *)d;
        else cudaGetLastError();
        if (!image_target && (hk == HostKind::mixed || (npix * sizeof(uint32_t) >= (256u << 10) && hk == HostKind::pageable)) && reserve_stage(ctx->stage_out, npix * sizeof(uint32_t)))
        {
            image_target = (uint32_t *)ctx->stage_out.dev; // pageable image: K3 writes the context's own mapped staging
            image_staged = true;
        }
        if (!image_target && hk == HostKind::mixed) return fail(ctx, VRT_CUDA_E_NOMEM, "the image buffer is only partly page-locked and no staging buffer could be allocated");
        if (!image_target)
            if (int rc = reserve(ctx, ctx->out_image, npix * sizeof(uint32_t))) return rc;
    }
    if (radiance)
    {
        pin_host(ctx, radiance, npix * sizeof(float) * 4);
        const HostKind hk = host_kind(radiance, npix * sizeof(float) * 4);
        void *d = nullptr;
        if (hk == HostKind::locked && cudaHostGetDevicePointer(&d, radiance, 0) == cudaSuccess) rad_target = (float *)d;
        else cudaGetLastError();
        if (hk == HostKind::mixed) rad_mixed = true; // (copied through pageable scratch below: a CUDA copy into it would be refused)
        if (!rad_target)
            if (int rc = reserve(ctx, ctx->out_rad, npix * sizeof(float) * 4)) return rc;
    }
    if (running && !*running) return VRT_CUDA_INTERRUPTED;
    // with a `running` flag the statistics are read after the poll loop, not inside render_device (which would block)
    int rc = vrt_cuda_render_device(ctx, frame, image ? (image_target ? image_target : (uint32_t *)ctx->out_image.p) : nullptr,
                                    radiance ? (rad_target ? rad_target : (float *)ctx->out_rad.p) : nullptr, running ? nullptr : stats);
    if (rc) return rc;
    bool interrupted = false;
    if (running)
    {
        while (cudaEventQuery(ctx->ev[3]) == cudaErrorNotReady)
        {
            if (!*running && !interrupted)
            {
                if (int ra = vrt_cuda_abort(ctx, 1)) return fail(ctx, ra, "vrt_cuda_abort failed");
                interrupted = true;
            }
            std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
        cudaGetLastError();
        CU(cudaStreamSynchronize(ctx->stream));
        if (interrupted)
            if (int ra = vrt_cuda_abort(ctx, 0)) return fail(ctx, ra, "vrt_cuda_abort failed");
        if (stats)
            if (int rs = read_stats(ctx, stats)) return rs;
        if (interrupted) return VRT_CUDA_INTERRUPTED;
    }
    const FrameGeom &G = ctx->geom;
    const size_t row0 = (size_t)G.row_begin, rows = (size_t)(G.row_end - G.row_begin);
    if (image && !image_target) CU(cudaMemcpyAsync(image + row0 * G.W, (uint32_t *)ctx->out_image.p + row0 * G.W, rows * G.W * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (radiance && !rad_target)
    {
        float *dst = radiance + row0 * G.W * 4;
        if (rad_mixed)
        {
            if (!reserve_stage(ctx->stage_rad, rows * G.W * sizeof(float) * 4)) return fail(ctx, VRT_CUDA_E_NOMEM, "the radiance buffer is only partly page-locked and no staging buffer could be allocated");
            dst = (float *)ctx->stage_rad.p;
        }
        CU(cudaMemcpyAsync(dst, (float *)ctx->out_rad.p + row0 * G.W * 4, rows * G.W * sizeof(float) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    if (radiance && !rad_target && rad_mixed) std::memcpy(radiance + row0 * G.W * 4, ctx->stage_rad.p, rows * G.W * sizeof(float) * 4);
    if (image_staged) parallel_copy(image + row0 * G.W, (const uint32_t *)ctx->stage_out.p + row0 * G.W, rows * G.W * sizeof(uint32_t));
    return 0;
}

int vrt_cuda_render(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame, uint32_t *image, float *radiance, vrt_cuda_stats *stats)
{
    return render_host(ctx, frame, image, radiance, stats, nullptr);
}

int vrt_cuda_render_interruptible(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame, uint32_t *image, float *radiance, vrt_cuda_stats *stats,
                                  const volatile unsigned char *running)
{
    return render_host(ctx, frame, image, radiance, stats, running);
}

int vrt_cuda_frame_render(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame, uint32_t *image, float *radiance, vrt_cuda_stats *stats)
{
    if (int rc = vrt_cuda_tile(ctx, frame)) return rc;
    return vrt_cuda_render(ctx, frame, image, radiance, stats);
}
} // extern "C"
