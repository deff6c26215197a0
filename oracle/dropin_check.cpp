// oracle/dropin_check.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Proves the drop-in claim with the reference's OWN types: this translation unit includes the unmodified reference
// headers (<vrt/vrt.h>, compiled in place from /root/reference/src) next to include/vrt_cuda.hpp, builds BASELINE
// config 1 exactly as src/volumetric-ray-tracer/main.cpp:196-255 does, and renders it
//   (a) with the reference's CPU entries   vrt::tile_gaussians + vrt::[simd_]render_image          (rt.cpp:29, rt.h:227-404)
//   (b) with the CUDA entries of vrt_cuda.hpp, called with the very same camera_t / gaussians_t / tiles_t objects,
// then compares the packed images channel by channel.  Exit status 0 = every mode within 1 LSB per channel.
// Built by oracle/Makefile into oracle/_ref/dropin_check (links oracle/_ref/libvrt_ref_v3.so and libvrt_cuda.so);
// run by tests/test_gpu_parity.py::test_dropin_with_reference_types on the GPU box.
#include <vrt/vrt.h>

#include "vrt_cuda.hpp"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace vrt;

static int max_channel_diff(const u32 *a, const u32 *b, u64 n, int first_shift = 0)
{
    int worst = 0;
    for (u64 i = 0; i < n; ++i)
        for (int sh = first_shift; sh < 32; sh += 8)
        {
            const int d = std::abs((int)((a[i] >> sh) & 0xFF) - (int)((b[i] >> sh) & 0xFF));
            if (d > worst) worst = d;
        }
    return worst;
}

int main(int argc, char **argv)
{
    const u64 width = 256, height = 256, tiles_per_axis = 16, threads = argc > 1 ? strtoul(argv[1], nullptr, 10) : 4;
    // scene: `-g 4` (main.cpp:196-204)
    std::vector<gaussian_t> scene;
    const u8 grid_dim = 4;
    for (u8 i = 0; i < grid_dim; ++i)
        for (u8 j = 0; j < grid_dim; ++j)
            scene.push_back(gaussian_t{
                .albedo{1.f - (i * grid_dim + j) / (f32)(grid_dim * grid_dim), 0.f, 0.f + (i * grid_dim + j) / (f32)(grid_dim * grid_dim), 1.f},
                .mu{-1.f + 1.f / grid_dim + i * 1.f / (grid_dim / 2.f), -1.f + 1.f / grid_dim + j * 1.f / (grid_dim / 2.f), 1.f},
                .sigma = 1.f / (2 * grid_dim),
                .magnitude = 1.f});
    gaussians_t gaussians{.gaussians = scene, .soa_gaussians = gaussian_vec_t::from_gaussians(scene)};
    // camera (main.cpp:248-255, rotation 0)
    vec4f_t origin{0.f, 0.f, -4.f};
    camera_t cam(origin.to_glm(), glm::vec3(0.f, 1.f, 0.f), glm::vec3(0.f, 0.f, 1.f), -90.f, 0.f, width, height, 1.f);
    const bool running = true;

    u32 *cpu = (u32 *)simd::aligned_malloc(sizeof(u32) * width * height);
    u32 *gpu = (u32 *)simd::aligned_malloc(sizeof(u32) * width * height);
    int failures = 0;
    auto report = [&](const char *what, int diff, int limit) {
        std::printf("%-58s max channel diff %d (limit %d) %s\n", what, diff, limit, diff <= limit ? "ok" : "FAIL");
        if (diff > limit) ++failures;
    };

    tiles_t tiles = tile_gaussians(2.f / tiles_per_axis, 2.f / tiles_per_axis, scene, cam.view_matrix);

    // mode 8: tiled, SIMD over pixels (the app default)
    simd_render_image(width, height, cpu, cam, origin, tiles, running, threads);
    bool r = cuda_simd_render_image(width, height, gpu, cam, origin, tiles, running, threads);
    report("mode 8  simd_render_image(tiles_t)    vs cuda_simd_render_image", max_channel_diff(cpu, gpu, width * height), 1);
    if (r) { std::printf("unexpected abort flag\n"); ++failures; }

    // mode 5: tiled, scalar exact erf
    render_image<radiance<transmittance>>(width, height, cpu, cam, origin, tiles, running, threads);
    cuda_render_image(width, height, gpu, cam, origin, tiles, running, threads);
    report("mode 5  render_image(tiles_t)         vs cuda_render_image", max_channel_diff(cpu, gpu, width * height), 1);

    // mode 4: untiled, SIMD over pixels
    simd_render_image(width, height, cpu, cam, origin, gaussians, running);
    cuda_simd_render_image(width, height, gpu, cam, origin, gaussians, running);
    report("mode 4  simd_render_image(gaussians_t) vs cuda_simd_render_image", max_channel_diff(cpu, gpu, width * height), 1);

    // device-side tiling: cuda_tile_gaussians + cuda_render_frame against the mode-8 CPU image
    simd_render_image(width, height, cpu, cam, origin, tiles, running, threads);
    const cuda_tiles_t dtiles = cuda_tile_gaussians(2.f / tiles_per_axis, 2.f / tiles_per_axis, scene);
    vrt_cuda_stats st;
    cuda_render_frame(width, height, gpu, cam, origin, dtiles, running, VRT_CUDA_MODE8, &st);
    report("mode 8  tile_gaussians+simd_render_image vs cuda_tile_gaussians+cuda_render_frame", max_channel_diff(cpu, gpu, width * height), 1);
    u64 ref_entries = 0;
    for (const gaussians_t &t : tiles.gaussians) ref_entries += t.gaussians.size();
    std::printf("list entries: reference %llu, device %llu %s\n", (unsigned long long)ref_entries, (unsigned long long)st.list_entries,
                ref_entries == st.list_entries ? "ok" : "FAIL");
    if (ref_entries != st.list_entries) ++failures;

    // the <Exp, Erf> template arguments of tests/img-error.cpp:42-43 as flags: fast_exp + A&S erf, and the Taylor erf
    simd_render_image<approx::simd_fast_exp, approx::simd_abramowitz_stegun_erf>(width, height, cpu, cam, origin, tiles, running, threads);
    cuda_simd_render_image(width, height, gpu, cam, origin, tiles, running, threads, cuda::approx_flags(VRT_CUDA_MODE8, cuda::exp_fn::fast, cuda::erf_fn::as));
    report("simd_render_image<simd_fast_exp, simd_abramowitz_stegun_erf> vs approx_flags(fast, as)", max_channel_diff(cpu, gpu, width * height), 2);
    simd_render_image<approx::vcl_exp<sizeof(simd::Vec<simd::Float>)>, approx::simd_taylor_erf>(width, height, cpu, cam, origin, tiles, running, threads);
    cuda_simd_render_image(width, height, gpu, cam, origin, tiles, running, threads, cuda::approx_flags(VRT_CUDA_MODE8, cuda::exp_fn::exact, cuda::erf_fn::taylor));
    report("simd_render_image<vcl_exp, simd_taylor_erf>               vs approx_flags(exact, taylor)", max_channel_diff(cpu, gpu, width * height), 2);

    // the `running` convention: a render that starts with running == false reports "interrupted"
    const bool stopped = false;
    if (!cuda_simd_render_image(width, height, gpu, cam, origin, tiles, stopped, threads)) { std::printf("running=false must return true\n"); ++failures; }

    simd::aligned_free(cpu);
    simd::aligned_free(gpu);
    delete gaussians.soa_gaussians;
    std::printf("dropin_check: %s\n", failures ? "FAILED" : "PASSED");
    return failures ? 1 : 0;
}
