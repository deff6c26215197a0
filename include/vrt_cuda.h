/* vrt_cuda.h -- C ABI of the B200-native (sm_100a) render path of the `vrt` Gaussian ray tracer.
 *
 * The reference has no FFI for this path: callers instantiate the header templates of src/vrt/rt.h
 * and link libvrt.so for tile_gaussians (src/vrt/rt.cpp).  This header is the boundary a maintainer
 * binds instead; include/vrt_cuda.hpp layers the reference's own C++ call shapes
 * (vrt::render_image / vrt::simd_render_image / vrt::tile_gaussians) on top of it.
 * Every entry point names the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a negative
 * VRT_CUDA_E_* code on failure (vrt_cuda_last_error() gives the text); nothing throws, exits or
 * falls back to a CPU implementation -- without a CUDA device vrt_cuda_create() fails.
 * One context drives one GPU and owns one stream; contexts are independent of each other, also on the SAME GPU: a frame's
 * geometry travels in the kernels' parameter blocks and the lists live in the context's own buffers, so two contexts can
 * have frames in flight on one device at the same time (the reference's entries are re-entrant, rt.h:259, 293).
 * A single context is driven by one host thread at a time; vrt_cuda_abort may be called from any thread.
 * Limits: images up to 65536 x 65536 with at most 2^22 8x4-pixel cells (about 134 Mpixel) per frame, reference tiles
 * per axis <= 1024, fewer than 2^31 Gaussians (device indices are 32-bit).
 */
#ifndef VRT_CUDA_H
#define VRT_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRT_CUDA_ABI_VERSION 5

/* Error codes */
#define VRT_CUDA_OK 0
#define VRT_CUDA_E_INVALID (-1)  /* bad argument / unsupported geometry          */
#define VRT_CUDA_E_CUDA (-2)     /* CUDA runtime error (text in last_error)      */
#define VRT_CUDA_E_STATE (-3)    /* call order: no Gaussians / no lists yet      */
#define VRT_CUDA_E_NOMEM (-4)
#define VRT_CUDA_INTERRUPTED 1   /* not an error: `running` went false mid-frame (the `true` of rt.h:244-246, 308-309) */

/* ---- frame flags ---------------------------------------------------------------------------- */
/* erf variant.  AS = Abramowitz-Stegun 7.1.27 with the reference's coefficients
 * (src/vrt/approx.cpp:90-110; what simd::erf resolves to, approx.h:110-118: modes 2-4, 6-8).
 * EXACT = libm-erff-class accuracy (template default of the scalar path, src/vrt/rt.h:32: modes 1, 5). */
#define VRT_CUDA_ERF_AS 0u
#define VRT_CUDA_ERF_EXACT 1u
#define VRT_CUDA_ERF_MASK 1u

/* Which Gaussians a pixel's list holds.
 * REFERENCE      : exactly the membership of vrt::tile_gaussians (src/vrt/rt.cpp:35-62) for the
 *                  pixel's reference tile -- the tiled modes 5-8, literal work.
 * REFERENCE_BOUND: that membership intersected with a conservative bound: Gaussians farther than
 *                  bound_sigmas * sigma from every ray of the pixel's 16x16 block are dropped (their
 *                  weight is < exp(-bound_sigmas^2 / 2)); same image within 1e-6, far less work.
 * ALL            : every Gaussian for every pixel -- the untiled modes 1-4 (src/vrt/rt.h:227-247, 315-337).
 * BOUND          : ALL intersected with the per-block bound. */
#define VRT_CUDA_LIST_REFERENCE (0u << 2)
#define VRT_CUDA_LIST_REFERENCE_BOUND (1u << 2)
#define VRT_CUDA_LIST_ALL (2u << 2)
#define VRT_CUDA_LIST_BOUND (3u << 2)
#define VRT_CUDA_LIST_MASK (3u << 2)

/* 8-bit quantisation: TRUNCATE = (u32)(min(c,1)*255) of the scalar entry points (rt.h:238-243, 278-283);
 * NEAREST = round-to-nearest-even of simd::cvts in the SIMD entry points (rt.h:329-333, 373-377). */
#define VRT_CUDA_QUANT_TRUNCATE (0u << 4)
#define VRT_CUDA_QUANT_NEAREST (1u << 4)
/* Alpha byte: 0xFF everywhere except the tiled SIMD entry (mode 8), which stores
 * min(1, sum albedo.w * inner) * 255 (rt.h:373, 377). */
#define VRT_CUDA_ALPHA_OPAQUE (0u << 5)
#define VRT_CUDA_ALPHA_FROM_W (1u << 5)
/* Evaluate the lists literally.  By default an entry whose weight exp(-d^2 / 2 sigma^2) is exactly 0 in fp32 for every ray
 * of an 8x4-pixel cell (farther than ~13.2 sigma from all of them) is dropped from that cell: in the literal list modes
 * (REFERENCE, ALL, caller-supplied tiles_t) vrt_cuda_tile / vrt_cuda_set_tile_lists intersect the lists with "visible from
 * the cell" once per frame, and K2 skips warp-uniformly what is left.  Such entries add exactly 0 to every sum, so the image
 * is the literal one up to the order of the fp32 additions (differences ~1e-7); vrt_cuda_stats.terms_listed,
 * vrt_cuda_get_lists and the membership they report stay literal.  NO_SKIP (given to the tile call AND the render call)
 * keeps every entry and evaluates every listed term: terms_executed == terms_listed. */
#define VRT_CUDA_NO_SKIP (1u << 6)
/* Banded evaluation -- the DEFAULT wherever K2 walks depth-sorted per-cell lists (every *_BOUND mode and the visible lists of
 * the literal modes).  K1 sorts every cell's list by depth along the cell's centre ray; an occluder that lies >= t_sat
 * standard widths in front of (behind) every sample of an emitter pair, for every pixel of the cell, is resolved as the
 * constant +A (-A) it evaluates to -- erf is saturated to +-1 in fp32 there -- and only the band of occluders around the
 * pair's own depth is evaluated term by term.  Same image (sums are reordered: differences ~1e-6), several times fewer
 * evaluated terms; vrt_cuda_stats.terms_saturated counts the terms resolved that way.  DEPTH_WINDOW is the round-1 name of
 * the opt-in and is still accepted (it now only asserts that the lists can take the banded kernel). */
#define VRT_CUDA_DEPTH_WINDOW (1u << 7)
/* Evaluate every listed term (the round-1 default): no saturation shortcut, no early exit.  terms_executed then counts
 * every (pixel, sample, occluder) triple the lists name, minus the warp-uniform zero-weight skips. */
#define VRT_CUDA_EVAL_ALL (1u << 12)
/* Transmittance early exit (banded kernel).  ln T(s) is non-increasing in s (every weight A_j >= 0, erf increasing), so once
 * T at a sample no remaining emitter's samples precede is below eps / (remaining emission weight) for every pixel of the
 * cell, the emitters behind change no channel by more than eps = 1e-6 and the cell stops (a warp-uniform break; replaces the
 * unconditional emitter loop rt.h:209-221).  vrt_cuda_stats.terms_terminated counts the terms dropped.  Disabled
 * automatically when the scene holds a negative magnitude or a non-positive sigma; NO_TERMINATE disables it explicitly. */
#define VRT_CUDA_NO_TERMINATE (1u << 13)

/* The reference's alternative approximations as selectable device functions (src/vrt/approx.h:10-46; the template
 * arguments <Exp, Erf> of the render entries, src/vrt/rt.h:315, 344, exercised by tests/img-error.cpp:40-43).
 * An erf selection overrides the ERF_AS / ERF_EXACT bit.  Exp replaces the exponential of c_bar and of the final
 * exp(T) (rt.h:115, 126); the density at the sample points keeps the exact exponential, as in the reference
 * (rt.h:218 uses the pdf's template default).  These run on a plain kernel without the saturation / sign shortcuts
 * (the approximations are neither odd nor continuous) and cannot be combined with VRT_CUDA_DEPTH_WINDOW. */
#define VRT_CUDA_APPROX_ERF_SPLINE (1u << 8)         /* approx::spline_erf,        approx.cpp:9-23   */
#define VRT_CUDA_APPROX_ERF_SPLINE_MIRROR (2u << 8)  /* approx::spline_erf_mirror, approx.cpp:45-56  */
#define VRT_CUDA_APPROX_ERF_TAYLOR (3u << 8)         /* approx::taylor_erf,        approx.cpp:64-77  */
#define VRT_CUDA_APPROX_ERF_MASK (3u << 8)
#define VRT_CUDA_APPROX_EXP_FAST (1u << 10)          /* approx::simd_fast_exp,     approx.cpp:112-137 (range-clamped) */
#define VRT_CUDA_APPROX_EXP_SPLINE (2u << 10)        /* approx::spline_exp,        approx.cpp:141-163 */
#define VRT_CUDA_APPROX_EXP_MASK (3u << 10)

/* Function ids of vrt_cuda_approx_table(): the columns of tests/accuracy.cpp's erf.csv / exp.csv. */
#define VRT_CUDA_FN_SPLINE_ERF 0
#define VRT_CUDA_FN_SPLINE_ERF_MIRROR 1
#define VRT_CUDA_FN_TAYLOR_ERF 2
#define VRT_CUDA_FN_AS_ERF 3
#define VRT_CUDA_FN_ERF 4        /* the device's libm-class erf (VRT_CUDA_ERF_EXACT) */
#define VRT_CUDA_FN_EXP 5        /* the device's exp (MUFU.EX2)                      */
#define VRT_CUDA_FN_FAST_EXP 6
#define VRT_CUDA_FN_SPLINE_EXP 7

/* Flag sets reproducing the reference's modes (src/volumetric-ray-tracer/main.cpp:150-177). */
#define VRT_CUDA_MODE1 (VRT_CUDA_ERF_EXACT | VRT_CUDA_LIST_ALL | VRT_CUDA_QUANT_TRUNCATE | VRT_CUDA_ALPHA_OPAQUE)
#define VRT_CUDA_MODE4 (VRT_CUDA_ERF_AS | VRT_CUDA_LIST_ALL | VRT_CUDA_QUANT_NEAREST | VRT_CUDA_ALPHA_OPAQUE)
#define VRT_CUDA_MODE5 (VRT_CUDA_ERF_EXACT | VRT_CUDA_LIST_REFERENCE | VRT_CUDA_QUANT_TRUNCATE | VRT_CUDA_ALPHA_OPAQUE)
#define VRT_CUDA_MODE8 (VRT_CUDA_ERF_AS | VRT_CUDA_LIST_REFERENCE | VRT_CUDA_QUANT_NEAREST | VRT_CUDA_ALPHA_FROM_W)

/* One frame's camera and geometry: what main.cpp:263-297 passes to tile_gaussians + the render entry. */
typedef struct vrt_cuda_frame
{
    float view[16];        /* camera_t::view_matrix, column-major (src/vrt/camera.cpp:52)                 */
    float origin[4];       /* ray origin, the `origin` argument of render_image (rt.h:228)                */
    uint32_t width;        /* image width / height in pixels                                              */
    uint32_t height;
    uint32_t tiles_x;      /* reference tiles per axis (main.cpp --tiles, tw = 2/tiles_x); ignored for    */
    uint32_t tiles_y;      /*   LIST_ALL / LIST_BOUND.  width % tiles_x == 0 and height % tiles_y == 0.   */
    uint32_t flags;        /* VRT_CUDA_* above                                                            */
    float bound_sigmas;    /* k of the *_BOUND list modes; <= 0 selects the default 6.0                   */
    uint32_t row_begin;    /* render only pixel rows [row_begin, row_end); 0,0 = whole image.  Used for   */
    uint32_t row_end;      /*   multi-GPU row bands; must be multiples of 16 (or the image height).       */
} vrt_cuda_frame;

/* What a render did.  "terms" are pixel-Gaussian evaluations: one evaluation of
 * A_j * erf(s r_j - m_j) for one pixel, one sample point s and one occluder j (loop body rt.h:107-124). */
typedef struct vrt_cuda_stats
{
    uint64_t n_gaussians;     /* scene size                                                               */
    uint64_t n_cells;         /* lists built (reference tiles or 16x16 blocks)                            */
    uint64_t list_entries;    /* sum over cells of list length                                            */
    uint32_t max_list;        /* longest list                                                             */
    uint32_t n_launches;      /* kernels launched by the last tile+render                                 */
    double terms_listed;      /* sum over rendered pixels of 5 * n^2 for the list used (E of SURVEY 8(d)) */
    double terms_executed;    /* terms actually evaluated (<= terms_listed: warp-uniform zero-weight skip)*/
    float ms_tile;            /* device time of the cull/tile kernels (CUDA events)                       */
    float ms_render;          /* device time of the render kernel(s)                                      */
    float ms_total;           /* first kernel to last kernel / copy of the call                           */
    uint32_t slice;           /* emitters per work item of split cells used by this frame (see vrt_cuda_set_slice)  */
    double terms_saturated;   /* banded evaluation: terms resolved by the saturation shortcut (not in terms_executed) */
    double terms_terminated;  /* banded evaluation: terms dropped by the transmittance early exit                    */
} vrt_cuda_stats;

typedef struct vrt_cuda_ctx vrt_cuda_ctx;

/* Lifetime.  `device` is the CUDA ordinal (one context per GPU / per rank). */
int vrt_cuda_create(int device, vrt_cuda_ctx **ctx_out);
void vrt_cuda_destroy(vrt_cuda_ctx *ctx);
const char *vrt_cuda_last_error(const vrt_cuda_ctx *ctx); /* ctx may be NULL: error of a failed create */
int vrt_cuda_abi_version(void);

/* Scene upload: replaces gaussians_t / `staging_gaussians` handed to tile_gaussians and the render entries
 * (main.cpp:209, 261-263).  `aos` is n records of vrt::gaussian_t (40 B each, src/vrt/types.h:195-200) in
 * HOST memory; the _device variant takes a device pointer on the context's GPU (e.g. the receive buffer of
 * an NCCL broadcast) and copies it device-to-device on the context's stream. */
int vrt_cuda_set_gaussians(vrt_cuda_ctx *ctx, const float *aos, uint64_t n);
int vrt_cuda_set_gaussians_device(vrt_cuda_ctx *ctx, const float *aos_dev, uint64_t n);

/* Tile projection + culling: replaces vrt::tile_gaussians(tw, th, gaussians, view) (src/vrt/rt.cpp:29-69)
 * with tw = 2/tiles_x, th = 2/tiles_y.  Builds the per-cell lists for `frame` on the device. */
int vrt_cuda_tile(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame);

/* Caller-supplied lists: the drop-in for the `const tiles_t &tiles` argument of the tiled render entries
 * (rt.h:252, 345).  `aos_concat` holds the tiles' gaussian_t records back to back (tiles_t::gaussians[t]
 * .gaussians, row-major tiles, y outer), tile t owning records [offsets[t], offsets[t+1]).  n_tiles must be
 * frame->tiles_x * frame->tiles_y.  The frame's list mode is ignored; the lists are used as given. */
int vrt_cuda_set_tile_lists(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame, const float *aos_concat,
                            const uint64_t *offsets, uint64_t n_tiles);

/* Read back the lists of the last vrt_cuda_tile() for membership parity with tile_gaussians: counts_out[c]
 * = length of cell c (row-major, y outer), idx_out = concatenated Gaussian indices in list order.  Either
 * may be NULL.  *n_cells_out / *n_entries_out receive the sizes needed. */
int vrt_cuda_get_lists(vrt_cuda_ctx *ctx, uint32_t *counts_out, uint64_t counts_cap, uint32_t *idx_out,
                       uint64_t idx_cap, uint64_t *n_cells_out, uint64_t *n_entries_out);

/* Render with the current lists: replaces the four render entries
 *   vrt::render_image<Radiance>(w, h, image, cam, origin, gaussians, running)            rt.h:227-247
 *   vrt::render_image<Radiance>(w, h, image, cam, origin, tiles, running, tc)            rt.h:251-310
 *   vrt::simd_render_image<Exp,Erf>(w, h, image, cam, origin, gaussians, running)        rt.h:315-337
 *   vrt::simd_render_image<Exp,Erf>(w, h, image, cam, origin, tiles, running, tc)        rt.h:344-404
 * `image` receives width*height packed 0xAARRGGBB pixels (row-major); `radiance` (may be NULL) receives
 * width*height float4 (x,y,z,w of the reference's vec4f_t colour before clamping) for parity tests.
 * Rows outside [row_begin,row_end) are left untouched.  Host pointers; the copy back is part of the call. */
int vrt_cuda_render(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame, uint32_t *image, float *radiance,
                    vrt_cuda_stats *stats);

/* Same with DEVICE output pointers (full-image sized buffers on the context's GPU); no copy back and no
 * host synchronisation unless `stats` is non-NULL.  Work is enqueued on the context's stream;
 * vrt_cuda_sync() waits for it.  Used by the multi-GPU path (bands are gathered with NCCL afterwards). */
int vrt_cuda_render_device(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame, uint32_t *image_dev, float *radiance_dev,
                           vrt_cuda_stats *stats);

/* The `const bool &running` argument of the reference's entries (rt.h:228, 252, 316, 346): `running` points at the caller's
 * flag (one byte, C++ bool), which another thread may clear while the frame renders (main.cpp:244: the viewer thread).  The
 * reference polls it per pixel / per tile (rt.h:244-246, 289, 334, 382) and returns true; here the host polls it while the
 * frame is in flight and raises a word in device memory the persistent render warps check before every work item.  Returns
 * VRT_CUDA_INTERRUPTED (the image is then partial and not copied back), 0 when the frame completed, < 0 on error. */
int vrt_cuda_render_interruptible(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame, uint32_t *image, float *radiance,
                                  vrt_cuda_stats *stats, const volatile unsigned char *running);
/* Raise (1) / clear (0) the abort word directly, from any thread: a vrt_cuda_render / vrt_cuda_render_device in flight on
 * this context stops taking work items.  The caller clears it before the next frame. */
int vrt_cuda_abort(vrt_cuda_ctx *ctx, int on);

/* Caller buffers.  The reference's callers hand in plain (pageable) memory -- `image` is simd::aligned_malloc'd once and
 * reused every frame (main.cpp:245, 338).  Such buffers go through page-locked staging that the context owns: the render
 * kernel writes the frame into a mapped staging image while it is computed and host threads copy it to `image`; the scene is
 * copied into staging by the same threads and uploaded chunk by chunk behind them (4 ms per frame on BASELINE config 5
 * against a page-locked caller buffer; a pageable cudaMemcpy, staged by the driver, cost 7 ms).
 * With pinning on, vrt_cuda_set_gaussians and vrt_cuda_render page-lock the buffer they are given instead (cudaHostRegister,
 * once per pointer; up to four registrations are kept): the scene is read and the image written in place, no host copy
 * remains.  OPT-IN, because the caller must keep such a buffer alive until it turns pinning off again (which releases every
 * registration) or destroys the context.  A buffer that lies only partly inside a registration (small heap buffers sharing a
 * page) is treated like a pageable one. */
int vrt_cuda_set_host_pinning(vrt_cuda_ctx *ctx, int on);
/* Explicit form: page-lock [p, p + bytes) (rounded out to pages) for every device of the process until vrt_cuda_unpin_buffer(p);
 * what an application that owns its image buffer for its whole run calls once (the host app does). */
int vrt_cuda_pin_buffer(vrt_cuda_ctx *ctx, void *p, uint64_t bytes);
int vrt_cuda_unpin_buffer(vrt_cuda_ctx *ctx, void *p);

/* Multi-GPU output without a gather step (one process per GPU).  The rank that owns the frame allocates the image on its GPU
 * with vrt_cuda_peer_image_create and hands the 64-byte handle to the other processes (any byte transport); they map it with
 * vrt_cuda_peer_image_open and pass the mapped pointer as `image_dev` of vrt_cuda_render_device.  K3's 16-byte stores of a rank's
 * row band then go straight into the owner's memory over NVLink / NVSwitch while the band is still being computed: the transfer
 * is part of the render kernel, and what remains of the exchange is one barrier.  (The reference composes its tiles into the
 * caller's image with a copy loop after each tile, rt.h:388-399; this is the same step across GPUs.)  _close unmaps an imported
 * image or frees an owned one.  An image can be opened once per importing process; not within the process that created it. */
#define VRT_CUDA_PEER_HANDLE_BYTES 64
int vrt_cuda_peer_image_create(vrt_cuda_ctx *ctx, uint64_t bytes, void **image_dev_out, unsigned char *handle_out);
int vrt_cuda_peer_image_open(vrt_cuda_ctx *ctx, const unsigned char *handle, void **image_dev_out);
int vrt_cuda_peer_image_close(vrt_cuda_ctx *ctx, void *image_dev);

/* vrt_cuda_tile + vrt_cuda_render in one call: one iteration of the app's frame loop (main.cpp:257-297). */
int vrt_cuda_frame_render(vrt_cuda_ctx *ctx, const vrt_cuda_frame *frame, uint32_t *image, float *radiance,
                          vrt_cuda_stats *stats);

/* Per-tile-row cost (sum over the row's cells of pixels * 5 * n^2) of the last vrt_cuda_tile(), for
 * work-balanced row bands (include/vrt_host.h: vrt_host_row_bands).  rows_out: tiles_y (or block rows). */
int vrt_cuda_row_costs(vrt_cuda_ctx *ctx, double *rows_out, uint32_t rows_cap, uint32_t *n_rows_out, uint32_t *row_height_px_out);

/* Kernel tuning knob for benchmarks (not part of the reference's surface): emitters per register block
 * Q in {2,4,6,8} and packed f32x2 arithmetic on (1) / off (0).  Defaults are the tuned values. */
int vrt_cuda_set_tuning(vrt_cuda_ctx *ctx, int emitter_block, int packed_f32x2);
/* Resident CTAs per SM of the banded kernel: 4 (default, 128 registers) or 5 (96 registers); a benchmarking knob. */
int vrt_cuda_set_band_tuning(vrt_cuda_ctx *ctx, int ctas_per_sm);
/* Further A/B knobs are environment variables read once by vrt_cuda_create (they select between kernels that produce the same
 * picture up to the order of the fp32 sums; the defaults are the measured best, profiles/r02_long_lists_ab.md):
 *   VRT_CUDA_LONG_BAND=0      lists of 153..832 entries go to k2_render's in-loop saturation test instead of k2_band_long
 *   VRT_CUDA_LONG_WIDE=<s>    K1 marks such a list "wide" (-> in-loop test) when its middle emitter must evaluate more than the
 *                             share s of the list (default 0.4; >= 1: never)
 *   VRT_CUDA_WIN_MINB=1       the in-loop test as one 8-warp CTA per SM (round 1) instead of three 4-warp CTAs
 *   VRT_CUDA_MAX_LIST_ENTRIES lowers the 2^32 list-entry limit (so that a test can reach VRT_CUDA_E_NOMEM) */

/* Work items of heavy cells.  A cell whose list is longer than 3 x slice entries is rendered as ceil(n / slice) independent
 * items (emitter ranges) whose partial radiances are summed in slice order; the fp32 result depends on the grouping, so two
 * renders are bit-identical only if they use the same slice (8, 16, 32, 64, 128 or 256).  slice = 0 (default) picks it per frame
 * from the listed work;
 * a multi-GPU caller that wants its gathered bands to equal a single-GPU frame bit for bit asks for the automatic choice
 * of the FULL frame at its share of the work (vrt_cuda_auto_slice after a full-frame vrt_cuda_tile, share = 1 / ranks) and
 * sets it on every rank.  A pinned slice also pins the emitter register block of the render kernel (8; otherwise chosen from
 * the mean list length of the rendered band), the only other choice that regroups the sums. */
int vrt_cuda_set_slice(vrt_cuda_ctx *ctx, int slice);
int vrt_cuda_auto_slice(vrt_cuda_ctx *ctx, double share, int *slice_out);

/* Roofline probe: FP32 FMA throughput of the context's GPU in TFLOP/s (FMA = 2 flops), measured with
 * independent FFMA chains (packed_f32x2 = 0) or FFMA2 chains (1).  bench.py reports it beside the nominal peak. */
int vrt_cuda_fp32_peak(vrt_cuda_ctx *ctx, int packed_f32x2, double *tflops_out);

/* Roofline probe: ceiling of K2's inner-term instruction mix (terms/s) with `pairs` (5, 10, 20) independent packed
 * pairs per thread and `ctas_per_sm` resident 256-thread CTAs -- no loads, setup or control flow. */
int vrt_cuda_term_peak(vrt_cuda_ctx *ctx, int pairs, int ctas_per_sm, double *terms_per_s_out);

/* Pipe-mix probe: chain steps/s where one step = nf FFMA2 + nm MUFU.RCP + nl LOP3 (a few fixed combinations). */
int vrt_cuda_mix_peak(vrt_cuda_ctx *ctx, int nf, int nm, int nl, double *steps_per_s_out);

/* y[i] = f(x[i]) for one of the VRT_CUDA_FN_* device functions, evaluated on the GPU (host pointers): the drop-in for the
 * tabulation loops of tests/accuracy.cpp:16-52, and the way the parity tests pin every device approximation. */
int vrt_cuda_approx_table(vrt_cuda_ctx *ctx, int fn, const float *x, float *y, uint64_t n);

/* Throughput of one VRT_CUDA_FN_* device function in values per second (independent argument chains, no memory traffic):
 * the GPU counterpart of tests/approx_cycles.cpp, which reports CPU cycles per value for the same functions. */
int vrt_cuda_approx_rate(vrt_cuda_ctx *ctx, int fn, double *values_per_s_out);

int vrt_cuda_sync(vrt_cuda_ctx *ctx);
/* The context's cudaStream_t as an integer (for ordering NCCL / torch work after a render_device). */
uint64_t vrt_cuda_stream(vrt_cuda_ctx *ctx);
int vrt_cuda_device(const vrt_cuda_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* VRT_CUDA_H */
