// vrt_cuda.hpp -- the reference's own C++ call shapes on top of the C ABI (include/vrt_cuda.h).
//
// Header-only and duck-typed: every overload is a template over the caller's camera / vector / scene types, so the
// same header compiles against the reference's structs (src/vrt/types.h, src/vrt/camera.h -- include <vrt/vrt.h>
// first) and against the stand-alone mirrors in include/vrt_types.hpp.  The only members used are the ones the
// reference's render entries read:
//     Camera   : .view_matrix  (glm::mat4-like, m[col][row])                          camera.h:26
//     Vec4     : .x .y .z .w                                                           types.h:19-21
//     Gaussians: .gaussians    (std::vector of 40-byte gaussian_t)                     types.h:266-270
//     Tiles    : .gaussians    (std::vector<Gaussians>), .tw .th .w .h                 types.h:272-287
//
// Replaces (argument lists identical up to a trailing, defaulted `flags`):
//     vrt::render_image<Radiance>(w, h, image, cam, origin, gaussians, running)            rt.h:227-247  -> cuda_render_image
//     vrt::render_image<Radiance>(w, h, image, cam, origin, tiles, running, tc)            rt.h:251-310  -> cuda_render_image
//     vrt::simd_render_image<Exp,Erf>(w, h, image, cam, origin, gaussians, running)        rt.h:315-337  -> cuda_simd_render_image
//     vrt::simd_render_image<Exp,Erf>(w, h, image, cam, origin, tiles, running, tc)        rt.h:344-404  -> cuda_simd_render_image
//     vrt::tile_gaussians(tw, th, gaussians, view)                                         rt.cpp:29-69  -> cuda_tile_gaussians
// Return convention as in the reference: true = the render was interrupted (`running` went false), false = completed
// (rt.h:244-246, 308-309).  Like the reference (definitions.h:23-30, gaussians-from-file.cpp:10-17) fatal conditions --
// here: any VRT_CUDA_E_* from the C ABI, e.g. no GPU -- print a message and terminate the process.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "vrt_cuda.h"

namespace vrt
{
namespace cuda
{
    /// Process-wide context for `device` (created on first use, destroyed at exit).
    inline vrt_cuda_ctx *context(int device = 0)
    {
        struct holder
        {
            vrt_cuda_ctx *ctx[16] = {};
            ~holder()
            {
                for (vrt_cuda_ctx *c : ctx)
                    if (c) vrt_cuda_destroy(c);
            }
        };
        static holder h;
        if (device < 0 || device >= 16)
        {
            std::fprintf(stderr, "[ ERROR ]\tvrt::cuda: device %d out of range\n", device);
            std::exit(EXIT_FAILURE);
        }
        if (!h.ctx[device] && vrt_cuda_create(device, &h.ctx[device]) != VRT_CUDA_OK)
        {
            std::fprintf(stderr, "[ ERROR ]\tvrt::cuda: %s\n", vrt_cuda_last_error(nullptr));
            std::exit(EXIT_FAILURE);
        }
        return h.ctx[device];
    }

    /// Opt in to page-locking the buffers handed to the render entries (`image`, the Gaussian vector): the reference's main
    /// allocates `image` once with simd::aligned_malloc and reuses it every frame (main.cpp:245, 338), so one registration
    /// makes every frame's copy-back run at PCIe rate.  The buffers must stay allocated until pin_host_buffers(false).
    inline void pin_host_buffers(bool on, int device = 0) { vrt_cuda_set_host_pinning(context(device), on ? 1 : 0); }

    inline void check(vrt_cuda_ctx *ctx, int rc, const char *what)
    {
        if (rc == VRT_CUDA_OK || rc == VRT_CUDA_INTERRUPTED) return;
        std::fprintf(stderr, "[ ERROR ]\tvrt::cuda: %s failed (%d): %s\n", what, rc, vrt_cuda_last_error(ctx));
        std::exit(EXIT_FAILURE);
    }

    template <class Camera, class Vec4>
    vrt_cuda_frame make_frame(uint32_t width, uint32_t height, const Camera &cam, const Vec4 &origin, uint32_t flags,
                              uint32_t tiles_x = 1, uint32_t tiles_y = 1, float bound_sigmas = 0.f)
    {
        vrt_cuda_frame f;
        std::memset(&f, 0, sizeof(f));
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r) f.view[4 * c + r] = cam.view_matrix[c][r];
        f.origin[0] = origin.x; f.origin[1] = origin.y; f.origin[2] = origin.z; f.origin[3] = origin.w;
        f.width = width; f.height = height;
        f.tiles_x = tiles_x; f.tiles_y = tiles_y;
        f.flags = flags;
        f.bound_sigmas = bound_sigmas;
        return f;
    }

    /// Tile (optionally) + render while watching the caller's `running` flag: the reference polls it per pixel / per tile
    /// (rt.h:244-246, 289, 334, 382) and returns true when it went false; here the C ABI polls it while the frame is in
    /// flight and stops the persistent render warps (vrt_cuda_render_interruptible).
    inline bool render_running(vrt_cuda_ctx *ctx, const vrt_cuda_frame &f, uint32_t *image, vrt_cuda_stats *stats, const bool &running, bool tile)
    {
        static_assert(sizeof(bool) == 1, "the C ABI reads `running` as one byte");
        if (tile) check(ctx, vrt_cuda_tile(ctx, &f), "vrt_cuda_tile");
        if (!running) return true;
        const int rc = vrt_cuda_render_interruptible(ctx, &f, image, nullptr, stats, reinterpret_cast<const volatile unsigned char *>(&running));
        check(ctx, rc, "vrt_cuda_render_interruptible");
        return rc == VRT_CUDA_INTERRUPTED || !running;
    }

    /// Untiled entries: every Gaussian for every pixel (or the bounded lists if `flags` says so).
    template <class Camera, class Vec4, class Gaussians>
    bool render_gaussians(uint32_t width, uint32_t height, uint32_t *image, const Camera &cam, const Vec4 &origin,
                          const Gaussians &gaussians, const bool &running, uint32_t flags, vrt_cuda_stats *stats = nullptr, int device = 0)
    {
        static_assert(sizeof(gaussians.gaussians[0]) == 40, "gaussian_t must be 10 packed floats (src/vrt/types.h:195-200)");
        if (!running) return true;
        vrt_cuda_ctx *ctx = context(device);
        const vrt_cuda_frame f = make_frame(width, height, cam, origin, flags);
        check(ctx, vrt_cuda_set_gaussians(ctx, reinterpret_cast<const float *>(gaussians.gaussians.data()), gaussians.gaussians.size()), "vrt_cuda_set_gaussians");
        return render_running(ctx, f, image, stats, running, true);
    }

    /// Tiled entries: the caller's tiles_t supplies one list per reference tile (row-major, y outer).
    template <class Camera, class Vec4, class Tiles>
    bool render_tiles(uint32_t width, uint32_t height, uint32_t *image, const Camera &cam, const Vec4 &origin, const Tiles &tiles,
                      const bool &running, uint32_t flags, vrt_cuda_stats *stats = nullptr, int device = 0)
    {
        if (!running) return true;
        vrt_cuda_ctx *ctx = context(device);
        const vrt_cuda_frame f = make_frame(width, height, cam, origin, flags, (uint32_t)tiles.w, (uint32_t)tiles.h);
        const uint64_t n_tiles = tiles.gaussians.size();
        std::vector<uint64_t> offsets(n_tiles + 1, 0);
        for (uint64_t t = 0; t < n_tiles; ++t) offsets[t + 1] = offsets[t] + tiles.gaussians[t].gaussians.size();
        std::vector<float> concat(offsets[n_tiles] * 10 + 10);
        for (uint64_t t = 0; t < n_tiles; ++t)
        {
            const auto &list = tiles.gaussians[t].gaussians;
            static_assert(sizeof(list[0]) == 40, "gaussian_t must be 10 packed floats (src/vrt/types.h:195-200)");
            if (!list.empty()) std::memcpy(concat.data() + offsets[t] * 10, reinterpret_cast<const float *>(list.data()), list.size() * 40);
        }
        check(ctx, vrt_cuda_set_tile_lists(ctx, &f, concat.data(), offsets.data(), n_tiles), "vrt_cuda_set_tile_lists");
        return render_running(ctx, f, image, stats, running, false);
    }
} // namespace cuda

// ---- the <Exp, Erf> template arguments ---------------------------------------------------------------------------------
// The reference selects its approximations at compile time (rt.h:32, 315, 344; tests/img-error.cpp:40-43).  Here they are
// frame flags: cuda::approx_flags(mode, exp, erf) rewrites a VRT_CUDA_MODE* set so that, for example,
//     simd_render_image<approx::simd_fast_exp, approx::simd_abramowitz_stegun_erf>(w, h, img, cam, origin, tiles, running, tc)
// becomes
//     cuda_simd_render_image(w, h, img, cam, origin, tiles, running, tc, cuda::approx_flags(VRT_CUDA_MODE8, cuda::exp_fn::fast, cuda::erf_fn::as)).
namespace cuda
{
    enum class exp_fn { exact, fast, spline };                        // expf / vcl_exp, fast_exp, spline_exp        (approx.h:35-46)
    enum class erf_fn { as, exact, spline, spline_mirror, taylor };   // A&S, erff, spline_erf, _mirror, taylor_erf  (approx.h:10-33)

    constexpr uint32_t approx_flags(uint32_t mode_flags, exp_fn e, erf_fn f)
    {
        uint32_t flags = mode_flags & ~(VRT_CUDA_ERF_MASK | VRT_CUDA_APPROX_ERF_MASK | VRT_CUDA_APPROX_EXP_MASK);
        flags |= e == exp_fn::fast ? VRT_CUDA_APPROX_EXP_FAST : (e == exp_fn::spline ? VRT_CUDA_APPROX_EXP_SPLINE : 0u);
        flags |= f == erf_fn::exact    ? VRT_CUDA_ERF_EXACT
                 : f == erf_fn::spline ? VRT_CUDA_APPROX_ERF_SPLINE
                 : f == erf_fn::spline_mirror ? VRT_CUDA_APPROX_ERF_SPLINE_MIRROR
                 : f == erf_fn::taylor ? VRT_CUDA_APPROX_ERF_TAYLOR
                                       : VRT_CUDA_ERF_AS;
        return flags;
    }
} // namespace cuda

// ---- scalar entries: libm-class exp/erf, truncating quantisation (modes 1 and 5) -------------------------------------

/// Drop-in for vrt::render_image<radiance<transmittance>>(w, h, image, cam, origin, gaussians, running)  (rt.h:227-247).
template <class Camera, class Vec4, class Gaussians>
bool cuda_render_image(const uint32_t width, const uint32_t height, uint32_t *image, const Camera &cam, const Vec4 &origin,
                       const Gaussians &gaussians, const bool &running = true, const uint32_t flags = VRT_CUDA_MODE1)
{
    return cuda::render_gaussians(width, height, image, cam, origin, gaussians, running, flags);
}

/// Drop-in for vrt::render_image<radiance<transmittance>>(w, h, image, cam, origin, tiles, running, tc)  (rt.h:251-310).
/// `tc` (CPU thread count) is accepted and ignored: the GPU grid replaces the thread pool.
template <class Camera, class Vec4, class Tiles>
bool cuda_render_image(const uint32_t width, const uint32_t height, uint32_t *image, const Camera &cam, const Vec4 &origin,
                       const Tiles &tiles, const bool &running, const uint64_t tc, const uint32_t flags = VRT_CUDA_MODE5)
{
    (void)tc;
    return cuda::render_tiles(width, height, image, cam, origin, tiles, running, flags);
}

// ---- SIMD-over-pixels entries: A&S erf, round-to-nearest quantisation (modes 4 and 8) --------------------------------

/// Drop-in for vrt::simd_render_image(w, h, image, cam, origin, gaussians, running)  (rt.h:315-337).
template <class Camera, class Vec4, class Gaussians>
bool cuda_simd_render_image(const uint32_t width, const uint32_t height, uint32_t *image, const Camera &cam, const Vec4 &origin,
                            const Gaussians &gaussians, const bool &running = true, const uint32_t flags = VRT_CUDA_MODE4)
{
    return cuda::render_gaussians(width, height, image, cam, origin, gaussians, running, flags);
}

/// Drop-in for vrt::simd_render_image(w, h, image, cam, origin, tiles, running, tc)  (rt.h:344-404), including its
/// alpha-from-colour.w quirk (rt.h:373, 377).  No `tile_width % SIMD_FLOATS` restriction (rt.h:350) applies.
template <class Camera, class Vec4, class Tiles>
bool cuda_simd_render_image(const uint32_t width, const uint32_t height, uint32_t *image, const Camera &cam, const Vec4 origin,
                            const Tiles &tiles, const bool &running, const uint64_t tc, const uint32_t flags = VRT_CUDA_MODE8)
{
    (void)tc;
    return cuda::render_tiles(width, height, image, cam, origin, tiles, running, flags);
}

// ---- tiling on the device ---------------------------------------------------------------------------------------------

/// What cuda_tile_gaussians returns: the lists stay on the GPU (no per-tile deep copies as in rt.cpp:60-64).
struct cuda_tiles_t
{
    float tw, th;
    uint64_t w, h;
    uint32_t list_mode; // VRT_CUDA_LIST_*
    float bound_sigmas;
    int device;
};

/// Device-side replacement for the pair
///     tiles = vrt::tile_gaussians(tw, th, gaussians, cam.view_matrix);          main.cpp:263 / rt.cpp:29-69
///     vrt::[simd_]render_image(w, h, image, cam, origin, tiles, running, tc);    main.cpp:269-281
/// that never brings the lists to the host: upload the scene, then render frames with cuda_render_frame().
template <class GaussianVector>
cuda_tiles_t cuda_tile_gaussians(const float tw, const float th, const GaussianVector &gaussians,
                                 const uint32_t list_mode = VRT_CUDA_LIST_REFERENCE, const float bound_sigmas = 0.f, const int device = 0)
{
    static_assert(sizeof(gaussians[0]) == 40, "gaussian_t must be 10 packed floats (src/vrt/types.h:195-200)");
    vrt_cuda_ctx *ctx = cuda::context(device);
    cuda::check(ctx, vrt_cuda_set_gaussians(ctx, reinterpret_cast<const float *>(gaussians.data()), gaussians.size()), "vrt_cuda_set_gaussians");
    // same tile counts as tiles_t: w = ceil(2/tw), h = ceil(2/th)  (types.h:280)
    const uint64_t w = (uint64_t)std::ceil(2.f / tw), h = (uint64_t)std::ceil(2.f / th);
    return cuda_tiles_t{tw, th, w, h, list_mode, bound_sigmas, device};
}

/// One frame with device-built lists.  `mode_flags` = VRT_CUDA_MODE5 / VRT_CUDA_MODE8 (its list bits are replaced by the
/// cuda_tiles_t's list mode).  Returns the reference's bool; fills `stats` when given.
template <class Camera, class Vec4>
bool cuda_render_frame(const uint32_t width, const uint32_t height, uint32_t *image, const Camera &cam, const Vec4 &origin,
                       const cuda_tiles_t &tiles, const bool &running, const uint32_t mode_flags = VRT_CUDA_MODE8, vrt_cuda_stats *stats = nullptr)
{
    if (!running) return true;
    vrt_cuda_ctx *ctx = cuda::context(tiles.device);
    const uint32_t flags = (mode_flags & ~VRT_CUDA_LIST_MASK) | tiles.list_mode;
    const vrt_cuda_frame f = cuda::make_frame(width, height, cam, origin, flags, (uint32_t)tiles.w, (uint32_t)tiles.h, tiles.bound_sigmas);
    return cuda::render_running(ctx, f, image, stats, running, true);
}
} // namespace vrt
