// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A headless extern "C" shim over the UNMODIFIED reference sources, compiled in
// place from /root/reference/src by oracle/Makefile into oracle/_ref/libvrt_ref_*.so.
// Nothing from the reference is copied into this repository; this file only
// *calls* the reference's public entry points:
//
//   vrt::tile_gaussians                         src/vrt/rt.cpp:29-69
//   vrt::transmittance<Exp,Erf>                 src/vrt/rt.h:32-54
//   vrt::radiance<Tr>                           src/vrt/rt.h:146-164
//   vrt::render_image / vrt::simd_render_image  src/vrt/rt.h:227-404
//   vrt::camera_t                               src/vrt/camera.cpp:7-71
//   vrt::approx::abramowitz_stegun_erf          src/vrt/approx.cpp:90-99
//   vrt::approx::{spline_erf, spline_erf_mirror, taylor_erf, fast_exp, spline_exp} and their simd_ forms
//                                               src/vrt/approx.h:10-46
//   read_from_obj                               src/vrt/gaussians-from-file.cpp:7-44
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load the resulting library.
#include <vrt/vrt.h>
#include <vrt/gaussians-from-file.h>

#include <cstring>
#include <ctime>
#include <vector>

using namespace vrt;

namespace
{
    std::vector<gaussian_t> from_aos(const float *aos, u64 n)
    {
        static_assert(sizeof(gaussian_t) == 40, "gaussian_t must be 10 packed floats");
        std::vector<gaussian_t> g(n);
        if (n) std::memcpy((void *)g.data(), aos, n * sizeof(gaussian_t));
        return g;
    }

    // Camera exactly as src/volumetric-ray-tracer/main.cpp:248-255 builds it.
    struct app_camera_t
    {
        vec4f_t origin;
        camera_t cam;
        app_camera_t(f32 camera_offset, f32 focal, f32 initial_rot, u64 w, u64 h)
            : origin{0.f, 0.f, camera_offset},
              cam(origin.to_glm(), glm::vec3(0.f, 1.f, 0.f), glm::vec3(0.f, 0.f, 1.f), -90.f, 0.f, w, h, focal)
        {
            f32 angle = -90.f;
            cam.position = glm::vec3(glm::rotate(glm::mat4(1.f), glm::radians(initial_rot), glm::vec3(0.f, 1.f, 0.f)) * glm::vec4(cam.position, 1.f));
            origin = vec4f_t::from_glm(glm::vec4(cam.position, 0.f));
            angle -= initial_rot;
            cam.turn(angle, 0.f);
        }
    };

    f64 now_ms()
    {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }

    // Variant code = erf id | exp id << 4 (same coding as oracle/vrt_oracle.c):
    //   erf 0 libm erff (scalar default, rt.h:32)  1 abramowitz_stegun_erf  2 spline_erf  3 spline_erf_mirror  4 taylor_erf
    //   exp 0 libm expf                            1 fast_exp               2 spline_exp         (src/vrt/approx.h:10-46)
    typedef vec4f_t (*rad_fn)(const vec4f_t, const vec4f_t, const gaussians_t &);
    typedef f32 (*tr_fn)(const vec4f_t, const vec4f_t, const f32, const gaussians_t &);
    template <f32_func_t Exp>
    rad_fn pick_radiance_erf(int erf_id)
    {
        switch (erf_id)
        {
        case 1: return radiance<transmittance<Exp, approx::abramowitz_stegun_erf>>;
        case 2: return radiance<transmittance<Exp, approx::spline_erf>>;
        case 3: return radiance<transmittance<Exp, approx::spline_erf_mirror>>;
        case 4: return radiance<transmittance<Exp, approx::taylor_erf>>;
        default: return radiance<transmittance<Exp, erff>>;
        }
    }
    rad_fn pick_radiance(int variant)
    {
        switch ((variant >> 4) & 15)
        {
        case 1: return pick_radiance_erf<approx::fast_exp>(variant & 15);
        case 2: return pick_radiance_erf<approx::spline_exp>(variant & 15);
        default: return pick_radiance_erf<expf>(variant & 15);
        }
    }
    template <f32_func_t Exp>
    tr_fn pick_transmittance_erf(int erf_id)
    {
        switch (erf_id)
        {
        case 1: return transmittance<Exp, approx::abramowitz_stegun_erf>;
        case 2: return transmittance<Exp, approx::spline_erf>;
        case 3: return transmittance<Exp, approx::spline_erf_mirror>;
        case 4: return transmittance<Exp, approx::taylor_erf>;
        default: return transmittance<Exp, erff>;
        }
    }
    tr_fn pick_transmittance(int variant)
    {
        switch ((variant >> 4) & 15)
        {
        case 1: return pick_transmittance_erf<approx::fast_exp>(variant & 15);
        case 2: return pick_transmittance_erf<approx::spline_exp>(variant & 15);
        default: return pick_transmittance_erf<expf>(variant & 15);
        }
    }

    // tiled SIMD render entry (rt.h:344-404) with the SIMD forms of the same approximations; exact erf has no SIMD form
    // without SVML, so erf id 0 maps to simd::erf (= Abramowitz-Stegun, approx.h:110-118)
    typedef bool (*simd_render_fn)(const u32, const u32, u32 *, const camera_t &, const vec4f_t, const tiles_t &, const bool &, const u64);
    template <simd_f32_func_t Exp>
    simd_render_fn pick_simd_render_erf(int erf_id)
    {
        switch (erf_id)
        {
        case 2: return simd_render_image<Exp, approx::simd_spline_erf>;
        case 3: return simd_render_image<Exp, approx::simd_spline_erf_mirror>;
        case 4: return simd_render_image<Exp, approx::simd_taylor_erf>;
        default: return simd_render_image<Exp, approx::simd_abramowitz_stegun_erf>;
        }
    }
    simd_render_fn pick_simd_render(int variant)
    {
        switch ((variant >> 4) & 15)
        {
        case 1: return pick_simd_render_erf<approx::simd_fast_exp>(variant & 15);
        case 2: return pick_simd_render_erf<approx::simd_spline_exp>(variant & 15);
        default: return pick_simd_render_erf<approx::vcl_exp<sizeof(simd::Vec<simd::Float>)>>(variant & 15);
        }
    }
}

extern "C"
{
    int ref_simd_floats() { return (int)SIMD_FLOATS; }
    int ref_sizeof_gaussian() { return (int)sizeof(gaussian_t); }

    /// View matrix (column-major) + origin of the app camera (main.cpp:248-255).
    void ref_app_camera(float camera_offset, float focal, float initial_rot, uint64_t w, uint64_t h,
                        float *view16, float *origin4)
    {
        app_camera_t c(camera_offset, focal, initial_rot, w, h);
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i) view16[4 * j + i] = c.cam.view_matrix[j][i];
        origin4[0] = c.origin.x; origin4[1] = c.origin.y; origin4[2] = c.origin.z; origin4[3] = c.origin.w;
    }

    /// View matrix of a camera_t built from camera_create_info_t-style arguments (camera.h:8-18).
    void ref_camera_view(const float *pos3, float yaw, float pitch, float focal, uint64_t w, uint64_t h, float *view16)
    {
        camera_t cam(glm::vec3(pos3[0], pos3[1], pos3[2]), glm::vec3(0.f, 1.f, 0.f), glm::vec3(0.f, 0.f, 1.f), yaw, pitch, w, h, focal);
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i) view16[4 * j + i] = cam.view_matrix[j][i];
    }

    /// Unit ray directions for the given pixel ids, computed the way render_image does
    /// (rt.h:231-236: plane point - origin, vec4f_t::normalize).  dirs_out: n_pix x 4.
    void ref_pixel_dirs(const float *pos3, float yaw, float pitch, float focal, uint64_t w, uint64_t h,
                        const float *origin4, const uint64_t *pix, uint64_t n_pix, float *dirs_out)
    {
        camera_t cam(glm::vec3(pos3[0], pos3[1], pos3[2]), glm::vec3(0.f, 1.f, 0.f), glm::vec3(0.f, 0.f, 1.f), yaw, pitch, w, h, focal);
        const vec4f_t origin{origin4[0], origin4[1], origin4[2], origin4[3]};
        for (u64 k = 0; k < n_pix; ++k)
        {
            const u64 i = pix[k];
            vec4f_t dir = vec4f_t{.x = cam.projection_plane.xs[i], .y = cam.projection_plane.ys[i], .z = cam.projection_plane.zs[i]} - origin;
            dir.normalize();
            dirs_out[4 * k + 0] = dir.x; dirs_out[4 * k + 1] = dir.y; dirs_out[4 * k + 2] = dir.z; dirs_out[4 * k + 3] = dir.w;
        }
    }

    /// Tile membership of vrt::tile_gaussians.  The scene copy handed to the reference carries the
    /// Gaussian index in albedo.w (unused by the tiling code) so membership can be read back as
    /// indices.  counts_out: one entry per tile (row-major, y outer); idx_out: concatenated lists.
    /// n_lists = number of lists produced (= number of float-accumulated tile centres, rt.cpp:47-49).
    /// Returns the total number of list entries (may exceed idx_cap; then idx_out is truncated).
    uint64_t ref_tile_membership(float tw, float th, const float *aos, uint64_t n, const float *view16,
                                 uint64_t *tiles_w, uint64_t *tiles_h, uint64_t *n_lists, uint32_t *counts_out, uint64_t counts_cap,
                                 uint32_t *idx_out, uint64_t idx_cap)
    {
        std::vector<gaussian_t> g = from_aos(aos, n);
        for (u64 i = 0; i < n; ++i) g[i].albedo.w = (f32)i;
        glm::mat4 view;
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i) view[j][i] = view16[4 * j + i];
        tiles_t tiles = tile_gaussians(tw, th, g, view);
        *tiles_w = tiles.w;
        *tiles_h = tiles.h;
        *n_lists = tiles.gaussians.size();
        u64 total = 0;
        for (u64 t = 0; t < tiles.gaussians.size(); ++t)
        {
            const auto &list = tiles.gaussians[t].gaussians;
            if (t < counts_cap) counts_out[t] = (u32)list.size();
            for (const gaussian_t &e : list)
            {
                if (total < idx_cap) idx_out[total] = (u32)e.albedo.w;
                ++total;
            }
        }
        return tiles.gaussians.size() > counts_cap ? (u64)-1 : total;
    }

    /// Float radiance of the reference's SCALAR path for explicit rays and an explicit list:
    /// vrt::radiance<vrt::transmittance<expf, ERF>>(origin, dir, list)  (rt.h:146-164 / :32-54).
    void ref_radiance(const float *aos_list, uint64_t n_list, const float *origin4, const float *dirs, uint64_t n_rays,
                      int variant, float *rgba_out)
    {
        gaussians_t gs{.gaussians = from_aos(aos_list, n_list)};
        const vec4f_t origin{origin4[0], origin4[1], origin4[2], origin4[3]};
        rad_fn R = pick_radiance(variant);
        for (u64 k = 0; k < n_rays; ++k)
        {
            const vec4f_t dir{dirs[4 * k], dirs[4 * k + 1], dirs[4 * k + 2], dirs[4 * k + 3]};
            const vec4f_t c = R(origin, dir, gs);
            rgba_out[4 * k + 0] = c.x; rgba_out[4 * k + 1] = c.y; rgba_out[4 * k + 2] = c.z; rgba_out[4 * k + 3] = c.w;
        }
    }

    /// vrt::transmittance<expf, ERF>(o, n, s, list) for a sweep of s (tests/transmittance.cpp:24-32).
    void ref_transmittance(const float *aos_list, uint64_t n_list, const float *origin4, const float *dir4,
                           const float *s, uint64_t n_s, int variant, float *T_out)
    {
        gaussians_t gs{.gaussians = from_aos(aos_list, n_list)};
        const vec4f_t o{origin4[0], origin4[1], origin4[2], origin4[3]};
        const vec4f_t d{dir4[0], dir4[1], dir4[2], dir4[3]};
        for (u64 k = 0; k < n_s; ++k)
            T_out[k] = pick_transmittance(variant)(o, d, s[k], gs);
    }

    void ref_as_erf(const float *x, uint64_t n, float *y)
    {
        for (u64 i = 0; i < n; ++i) y[i] = approx::abramowitz_stegun_erf(x[i]);
    }

    /// The functions tests/accuracy.cpp tabulates (src/vrt/approx.h:10-46).  fn 0 spline_erf, 1 spline_erf_mirror, 2 taylor_erf,
    /// 3 abramowitz_stegun_erf, 4 erff, 5 expf, 6 fast_exp, 7 spline_exp; fn + 16 = the SIMD form of the same function
    /// (erff/expf have none: 4+16 -> simd::erf, 5+16 -> vcl_exp).  Returns 0, or -1 for an unknown fn.
    int ref_approx_table(int fn, const float *x, uint64_t n, float *y)
    {
        if (fn < 16)
        {
            f32_func_t f = nullptr;
            switch (fn)
            {
            case 0: f = approx::spline_erf; break;
            case 1: f = approx::spline_erf_mirror; break;
            case 2: f = approx::taylor_erf; break;
            case 3: f = approx::abramowitz_stegun_erf; break;
            case 4: f = erff; break;
            case 5: f = expf; break;
            case 6: f = approx::fast_exp; break;
            case 7: f = approx::spline_exp; break;
            default: return -1;
            }
            for (u64 i = 0; i < n; ++i) y[i] = f(x[i]);
            return 0;
        }
        simd_f32_func_t f = nullptr;
        switch (fn - 16)
        {
        case 0: f = approx::simd_spline_erf; break;
        case 1: f = approx::simd_spline_erf_mirror; break;
        case 2: f = approx::simd_taylor_erf; break;
        case 3: f = approx::simd_abramowitz_stegun_erf; break;
        case 4: f = simd::erf; break;
        case 5: f = approx::vcl_exp<sizeof(simd::Vec<simd::Float>)>; break;
        case 6: f = approx::simd_fast_exp; break;
        case 7: f = approx::simd_spline_exp; break;
        default: return -1;
        }
        for (u64 i = 0; i < n; i += SIMD_FLOATS)
        {
            f32 in[SIMD_FLOATS], out[SIMD_FLOATS];
            for (u64 k = 0; k < SIMD_FLOATS; ++k) in[k] = x[i + k < n ? i + k : n - 1];
            simd::storeu(out, f(simd::loadu(in)));
            for (u64 k = 0; k < SIMD_FLOATS && i + k < n; ++k) y[i + k] = out[k];
        }
        return 0;
    }

    /// The procedure of tests/img-error.cpp:27-43 for one SIMD variant: lists from tile_gaussians(tw, tw, scene, tiling_view),
    /// default camera_create_info_t camera (position 0, yaw -90, 256x256, focal 1), origin 0, tiled SIMD entry
    /// simd_render_image<Exp, Erf> (variant >= 0) or the scalar reference image render_image<radiance<transmittance>>
    /// (variant < 0, img-error.cpp:34).  image_out: w*h packed pixels.
    int ref_img_error_image(const float *aos, uint64_t n, float tw, const float *tiling_view16, uint64_t w, uint64_t h,
                            uint64_t threads, int variant, uint32_t *image_out)
    {
        std::vector<gaussian_t> g = from_aos(aos, n);
        glm::mat4 view;
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i) view[j][i] = tiling_view16[4 * j + i];
        tiles_t tiles = tile_gaussians(tw, tw, g, view);
        camera_create_info_t ci{};
        ci.width = w;
        ci.height = h;
        const camera_t cam(ci);
        const vec4f_t origin{0.f, 0.f, 0.f};
        u32 *image = (u32 *)simd::aligned_malloc(sizeof(u32) * w * h);
        const bool running = true;
        if (variant < 0) render_image<radiance<transmittance>>(w, h, image, cam, origin, tiles, running, threads);
        else pick_simd_render(variant)(w, h, image, cam, origin, tiles, running, threads);
        std::memcpy(image_out, image, sizeof(u32) * w * h);
        simd::aligned_free(image);
        return 0;
    }

    /// The app's frame (main.cpp:257-297) for one mode: tile_gaussians + the mode's render entry.
    /// mode 1..8 as in main.cpp:150-177.  times_ms_out = {tiling, draw}.  Returns list-term count
    /// sum_t P_t*5*n_t^2 (tiled modes) or W*H*5*N^2 (untiled) through terms_out.
    int ref_render_app(int mode, const float *aos, uint64_t n, uint64_t w, uint64_t h, uint64_t tiles_per_axis,
                       uint64_t threads, float camera_offset, float focal, float initial_rot,
                       uint32_t *image_out, double *times_ms_out, double *terms_out)
    {
        std::vector<gaussian_t> g = from_aos(aos, n);
        gaussians_t gaussians{.gaussians = g, .soa_gaussians = gaussian_vec_t::from_gaussians(g)};
        app_camera_t c(camera_offset, focal, initial_rot, w, h);
        u32 *image = (u32 *)simd::aligned_malloc(sizeof(u32) * w * h);
        const bool running = true;
        bool res = false;
        f64 t0 = now_ms();
        tiles_t tiles = tile_gaussians(2.f / tiles_per_axis, 2.f / tiles_per_axis, g, c.cam.view_matrix);
        f64 t1 = now_ms();
        switch (mode)
        {
        // the eight instantiations of main.cpp:269-294 (modes 2/6: the default Radiance = radiance<simd_transmittance>,
        // SIMD over occluders, rt.h:61-95; modes 3/7: simd_radiance, SIMD over emitters, rt.h:166-199)
        case 1: res = render_image<radiance<transmittance>>(w, h, image, c.cam, c.origin, gaussians, running); break;
        case 2: res = render_image(w, h, image, c.cam, c.origin, gaussians, running); break;
        case 3: res = render_image<simd_radiance>(w, h, image, c.cam, c.origin, gaussians, running); break;
        case 4: res = simd_render_image(w, h, image, c.cam, c.origin, gaussians, running); break;
        case 5: res = render_image<radiance<transmittance>>(w, h, image, c.cam, c.origin, tiles, running, threads); break;
        case 6: res = render_image(w, h, image, c.cam, c.origin, tiles, running, threads); break;
        case 7: res = render_image<simd_radiance>(w, h, image, c.cam, c.origin, tiles, running, threads); break;
        case 8: res = simd_render_image(w, h, image, c.cam, c.origin, tiles, running, threads); break;
        default: simd::aligned_free(image); delete gaussians.soa_gaussians; return -1;
        }
        f64 t2 = now_ms();
        times_ms_out[0] = t1 - t0;
        times_ms_out[1] = t2 - t1;
        f64 terms = 0.0;
        if (mode >= 5)
        {
            const f64 P = (f64)((u64)(w * tiles.tw / 2.f)) * (f64)((u64)(h * tiles.th / 2.f));
            for (const gaussians_t &t : tiles.gaussians) terms += P * 5.0 * (f64)t.gaussians.size() * (f64)t.gaussians.size();
        }
        else terms = (f64)w * h * 5.0 * (f64)n * (f64)n;
        *terms_out = terms;
        std::memcpy(image_out, image, sizeof(u32) * w * h);
        simd::aligned_free(image);
        delete gaussians.soa_gaussians;
        return res ? 1 : 0;
    }

    /// CPU baseline on explicit work: runs the reference's production render entry
    ///   vrt::simd_render_image(w', h', image, cam, origin, tiles, running, tc)   (rt.h:344-404, mode 8)
    /// over `n_tiles` tiles of tile_w x tile_h pixels laid out side by side (w' = n_tiles*tile_w,
    /// h' = tile_h).  The caller supplies each tile's Gaussian list (records, concatenated,
    /// list_offsets has n_tiles+1 entries) and the projection-plane point of every pixel of the
    /// w' x h' strip, so arbitrary tiles of a larger frame can be timed with exactly the lists the
    /// GPU used.  scalar_variant < 0: SIMD mode-8 entry; 0/1: scalar tiled entry (mode 5) with
    /// erff / A&S erf.  Returns draw time in ms; image_out is w'*h' packed pixels.
    double ref_render_tile_strip(const float *aos_concat, const uint64_t *list_offsets, uint64_t n_tiles,
                                 uint64_t tile_w, uint64_t tile_h, const float *plane_xs, const float *plane_ys,
                                 const float *plane_zs, const float *origin4, uint64_t threads, int scalar_variant,
                                 uint32_t *image_out)
    {
        const u64 w = n_tiles * tile_w, h = tile_h;
        std::vector<gaussians_t> lists(n_tiles);
        for (u64 t = 0; t < n_tiles; ++t)
        {
            lists[t].gaussians = from_aos(aos_concat + 10 * list_offsets[t], list_offsets[t + 1] - list_offsets[t]);
            lists[t].soa_gaussians = gaussian_vec_t::from_gaussians(lists[t].gaussians);
        }
        tiles_t tiles(lists, 2.f / (f32)n_tiles, 2.f); // tiles.w = n_tiles, tiles.h = 1
        camera_t cam(glm::vec3(0.f), glm::vec3(0.f, 1.f, 0.f), glm::vec3(0.f, 0.f, 1.f), -90.f, 0.f, w, h, 1.f);
        std::memcpy(cam.projection_plane.xs, plane_xs, sizeof(f32) * w * h);
        std::memcpy(cam.projection_plane.ys, plane_ys, sizeof(f32) * w * h);
        std::memcpy(cam.projection_plane.zs, plane_zs, sizeof(f32) * w * h);
        const vec4f_t origin{origin4[0], origin4[1], origin4[2], origin4[3]};
        u32 *image = (u32 *)simd::aligned_malloc(sizeof(u32) * w * h);
        const bool running = true;
        const f64 t0 = now_ms();
        if (scalar_variant < 0) simd_render_image(w, h, image, cam, origin, tiles, running, threads);
        else if (scalar_variant == 1) render_image<radiance<transmittance<expf, approx::abramowitz_stegun_erf>>>(w, h, image, cam, origin, tiles, running, threads);
        else render_image<radiance<transmittance<expf, erff>>>(w, h, image, cam, origin, tiles, running, threads);
        const f64 t1 = now_ms();
        std::memcpy(image_out, image, sizeof(u32) * w * h);
        simd::aligned_free(image);
        return t1 - t0;
    }

    /// read_from_obj (gaussians-from-file.cpp:7-44).  Returns the Gaussian count; copies up to cap.
    uint64_t ref_read_obj(const char *path, float *aos_out, uint64_t cap)
    {
        std::vector<gaussian_t> g = read_from_obj(path);
        const u64 m = g.size() < cap ? g.size() : cap;
        if (m) std::memcpy(aos_out, (const void *)g.data(), m * sizeof(gaussian_t));
        return g.size();
    }
}
