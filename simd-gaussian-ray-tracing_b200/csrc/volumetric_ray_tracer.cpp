// volumetric-ray-tracer (headless, CUDA) -- the caller of the hot path, restated from scratch.
//
// Keeps the command line of src/volumetric-ray-tracer/main.cpp:28-183 (-f/--file, -o/--output, -g/--grid[=dim],
// -w/--width, -h/--height, -t/--with-threads, -q/--quiet, --frames, --tiles, -r/--rotation, -i/--initial-rotation,
// -c/--camera-offset, --focal-length, -m/--mode) and its frame loop (main.cpp:257-334: tile, render, PNG name_<frame>.ext,
// `TIME:` / `AVG. TIME:` lines, camera orbit), but every frame is rendered by the sm_100a path behind include/vrt_cuda.h.
// Modes 1-8 select the same semantics as the reference's modes (erf variant, lists, quantisation, alpha); new:
//     -m 9   untiled semantics + k-sigma bounded lists   (mode 4 image, far less work)
//     -m 10  tiled semantics, reference lists AND k-sigma bound   (mode 8 image, far less work)   [default]
//     --gpus <n>      split every frame into n work-balanced row bands, one GPU each (in-process; bands land directly in
//                     the host image, no inter-GPU traffic is needed because bands do not overlap)
//     --bound <k>     k of the bounded list modes (default 6)
//     --synthetic <n> [--seed <s>] [--sigma-range <lo>,<hi>]   the frustum-filling random scene of BASELINE configs 4/5
// The Vulkan viewer (non-quiet mode) is out of scope: without -q the program warns and runs headless anyway.
#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <thread>
#include <vector>

#include "vrt_cuda.h"
#include "vrt_host.h"
#include "vrt_types.hpp"

static const char *const HELP_MSG =
    "volumetric-ray-tracer (CUDA, headless)\n"
    "  scene     -g [dim]              dim x dim Gaussian grid (default 4; wins over nothing, loses to -f)\n"
    "            -f <obj>              one Gaussian per OBJ vertex\n"
    "            --synthetic <n>       n random Gaussians filling the view frustum [--seed <s>] [--sigma-range <lo>,<hi> (log10)]\n"
    "  image     -w <px> / -h <px>     size (one of them sets both), --tiles <n> reference tiles per axis (16)\n"
    "            -o <file.png>         write the frame(s); <file>_<k>.png when --frames > 1\n"
    "  camera    -c <z> (-4), --focal-length <f> (1), -i <deg> start angle, -r <deg> total turn over --frames <n>\n"
    "  render    -m <mode>             1-8: the reference's modes (erf variant, lists, quantisation, alpha)\n"
    "                                  9: untiled semantics + k-sigma lists, 10: tiled semantics + k-sigma lists (default)\n"
    "            --bound <k> (6)       k of modes 9/10;   --gpus <n> row bands on n GPUs\n"
    "            --erf <fn> / --exp <fn>   swap in one of the library's other approximations (the <Exp, Erf> template arguments):\n"
    "                                  erf: as | exact | spline | spline-mirror | taylor     exp: exact | fast | spline\n"
    "  misc      -q (always headless), -t <n> (ignored: no thread pool), --help\n";

struct cmd_args_t
{
    uint64_t w = (uint64_t)-1, h = (uint64_t)-1, grid_dim = 4, thread_count = 1, nr_frames = 1, tiles = 16, mode = 10;
    uint64_t synthetic = 0, seed = 43, gpus = 1;
    uint32_t approx_set = 0, approx_clear = 0; // flag bits forced on / off by --erf / --exp
    char *outfile = nullptr, *infile = nullptr;
    bool use_grid = false, quiet = false;
    float rot = 360.f, initial_rot = 0.f, camera_offset = -4.f, focal_length = 1.f, bound = 6.f, sig_lo = -2.6f, sig_hi = -2.0f;

    cmd_args_t(int argc, char **argv)
    {
        static struct option opts[] = {{"grid", optional_argument, nullptr, 'g'},      {"file", required_argument, nullptr, 'f'},
                                       {"output", required_argument, nullptr, 'o'},    {"width", required_argument, nullptr, 'w'},
                                       {"height", required_argument, nullptr, 'h'},    {"with-threads", required_argument, nullptr, 't'},
                                       {"quiet", no_argument, nullptr, 'q'},           {"tiles", required_argument, nullptr, 'l'},
                                       {"mode", required_argument, nullptr, 'm'},      {"frames", required_argument, nullptr, 's'},
                                       {"rotation", required_argument, nullptr, 'r'},  {"initial-rotation", required_argument, nullptr, 'i'},
                                       {"camera-offset", required_argument, nullptr, 'c'}, {"focal-length", required_argument, nullptr, 0xfe},
                                       {"help", no_argument, nullptr, 0xff},           {"gpus", required_argument, nullptr, 0x100},
                                       {"bound", required_argument, nullptr, 0x101},   {"synthetic", required_argument, nullptr, 0x102},
                                       {"seed", required_argument, nullptr, 0x103},    {"sigma-range", required_argument, nullptr, 0x104},
                                       {"erf", required_argument, nullptr, 0x105},     {"exp", required_argument, nullptr, 0x106},
                                       {nullptr, 0, nullptr, 0}};
        int lidx;
        for (;;)
        {
            const int c = getopt_long(argc, argv, "r:m:qw:o:f:g:h:t:c:i:", opts, &lidx);
            if (c == -1) break;
            switch (c)
            {
            case 'g':
                use_grid = true;
                if (optarg != nullptr) grid_dim = strtoul(optarg, nullptr, 10);
                break;
            case 'f': infile = optarg; break;
            case 'o': outfile = optarg; break;
            case 'w':
                w = strtoul(optarg, nullptr, 10);
                if (h == (uint64_t)-1) h = w;
                break;
            case 'h':
                h = strtoul(optarg, nullptr, 10);
                if (w == (uint64_t)-1) w = h;
                break;
            case 't': thread_count = strtoul(optarg, nullptr, 10); break;
            case 'q': quiet = true; break;
            case 'l': tiles = strtoul(optarg, nullptr, 10); break;
            case 's': nr_frames = strtoul(optarg, nullptr, 10); break;
            case 'r': rot = strtof(optarg, nullptr); break;
            case 'i': initial_rot = strtof(optarg, nullptr); break;
            case 'c': camera_offset = strtof(optarg, nullptr); break;
            case 'm': mode = strtoul(optarg, nullptr, 10); break;
            case 0xfe: focal_length = strtof(optarg, nullptr); break;
            case 0xff: std::fputs(HELP_MSG, stdout); std::exit(EXIT_SUCCESS);
            case 0x100: gpus = strtoul(optarg, nullptr, 10); break;
            case 0x101: bound = strtof(optarg, nullptr); break;
            case 0x102: synthetic = strtoul(optarg, nullptr, 10); break;
            case 0x103: seed = strtoul(optarg, nullptr, 10); break;
            case 0x104: std::sscanf(optarg, "%f,%f", &sig_lo, &sig_hi); break;
            case 0x105:
            {
                const std::string v = optarg;
                approx_clear |= VRT_CUDA_ERF_MASK | VRT_CUDA_APPROX_ERF_MASK;
                approx_set &= ~(VRT_CUDA_ERF_MASK | VRT_CUDA_APPROX_ERF_MASK);
                if (v == "as") approx_set |= VRT_CUDA_ERF_AS;
                else if (v == "exact") approx_set |= VRT_CUDA_ERF_EXACT;
                else if (v == "spline") approx_set |= VRT_CUDA_APPROX_ERF_SPLINE;
                else if (v == "spline-mirror") approx_set |= VRT_CUDA_APPROX_ERF_SPLINE_MIRROR;
                else if (v == "taylor") approx_set |= VRT_CUDA_APPROX_ERF_TAYLOR;
                else { std::fputs(HELP_MSG, stderr); std::exit(EXIT_FAILURE); }
                break;
            }
            case 0x106:
            {
                const std::string v = optarg;
                approx_clear |= VRT_CUDA_APPROX_EXP_MASK;
                approx_set &= ~VRT_CUDA_APPROX_EXP_MASK;
                if (v == "fast") approx_set |= VRT_CUDA_APPROX_EXP_FAST;
                else if (v == "spline") approx_set |= VRT_CUDA_APPROX_EXP_SPLINE;
                else if (v != "exact") { std::fputs(HELP_MSG, stderr); std::exit(EXIT_FAILURE); }
                break;
            }
            default: std::fputs(HELP_MSG, stderr); std::exit(EXIT_FAILURE);
            }
        }
        if (w == (uint64_t)-1) w = 256;
        if (h == (uint64_t)-1) h = 256;
        if (use_grid && infile != nullptr) use_grid = false; // main.cpp:182
        if (nr_frames == 0) nr_frames = 1;
    }
};

/// flags reproducing the reference's mode switch (main.cpp:150-177, dispatch :269-294)
static uint32_t mode_flags(uint64_t mode)
{
    const uint32_t scalar_out = VRT_CUDA_QUANT_TRUNCATE | VRT_CUDA_ALPHA_OPAQUE;
    switch (mode)
    {
    case 1: return VRT_CUDA_MODE1;
    case 2: // simd_transmittance / simd_radiance inside the scalar render_image: A&S erf, truncating pack
    case 3: return VRT_CUDA_ERF_AS | VRT_CUDA_LIST_ALL | scalar_out;
    case 4: return VRT_CUDA_MODE4;
    case 5: return VRT_CUDA_MODE5;
    case 6:
    case 7: return VRT_CUDA_ERF_AS | VRT_CUDA_LIST_REFERENCE | scalar_out;
    case 8: return VRT_CUDA_MODE8;
    case 9: return (VRT_CUDA_MODE4 & ~VRT_CUDA_LIST_MASK) | VRT_CUDA_LIST_BOUND;
    default: return (VRT_CUDA_MODE8 & ~VRT_CUDA_LIST_MASK) | VRT_CUDA_LIST_REFERENCE_BOUND;
    }
}

/// the scene as the packed floats the C ABI takes (gaussian_t is 10 floats, types.h:195-200); nullptr for an empty scene
static float *aos(std::vector<vrt::gaussian_t> &g) { return g.empty() ? nullptr : &g[0].albedo.x; }

static double now_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

#define DIE(ctx, what)                                                                          \
    do                                                                                          \
    {                                                                                           \
        std::fprintf(stderr, "[ ERROR ]\t%s: %s\n", what, vrt_cuda_last_error(ctx));          \
        std::exit(EXIT_FAILURE);                                                                \
    } while (0)

int main(int argc, char **argv)
{
    cmd_args_t cmd(argc, argv);
    if (!cmd.quiet) std::fprintf(stderr, "[ WARNING ]\tthe Vulkan viewer is out of scope for the CUDA build; running headless as with --quiet\n");

    // ---- scene (main.cpp:188-204) ----
    std::vector<vrt::gaussian_t> gaussians;
    if (cmd.synthetic)
    {
        gaussians.resize(cmd.synthetic);
        vrt_host_scene_synthetic(cmd.synthetic, cmd.seed, cmd.sig_lo, cmd.sig_hi, aos(gaussians));
    }
    else if (cmd.infile != nullptr)
    {
        const uint64_t n = vrt_host_read_obj(cmd.infile, nullptr, 0);
        if (n == UINT64_MAX)
        {
            std::fprintf(stderr, "[ ERROR ]\tcannot read %s\n", cmd.infile);
            return EXIT_FAILURE;
        }
        gaussians.resize(n);
        vrt_host_read_obj(cmd.infile, aos(gaussians), n);
    }
    else
    {
        const uint32_t dim = (uint8_t)cmd.grid_dim; // u8 in the reference
        gaussians.resize((size_t)dim * dim);
        if (vrt_host_scene_grid(dim, aos(gaussians)) != (uint64_t)dim * dim)
        {
            std::fprintf(stderr, "[ ERROR ]\tbad grid dimension\n");
            return EXIT_FAILURE;
        }
    }

    // ---- devices ----
    const int n_gpus = cmd.gpus >= 1 && cmd.gpus <= 64 ? (int)cmd.gpus : 1;
    std::vector<vrt_cuda_ctx *> ctx(n_gpus, nullptr);
    for (int d = 0; d < n_gpus; ++d)
    {
        if (vrt_cuda_create(d, &ctx[d]) != VRT_CUDA_OK) DIE(nullptr, "vrt_cuda_create");
        if (vrt_cuda_set_gaussians(ctx[d], aos(gaussians), gaussians.size()) != VRT_CUDA_OK) DIE(ctx[d], "vrt_cuda_set_gaussians");
    }

    const uint32_t width = (uint32_t)cmd.w, height = (uint32_t)cmd.h;
    std::vector<uint32_t> image((size_t)width * height, 0u);
    // the frame buffer lives as long as the process (main.cpp:245, 338): page-lock it once so that every frame's copy-back
    // runs at PCIe rate (not fatal if the system refuses)
    const bool pinned = vrt_cuda_pin_buffer(ctx[0], image.data(), image.size() * sizeof(uint32_t)) == VRT_CUDA_OK;
    const uint32_t flags = (mode_flags(cmd.mode) & ~cmd.approx_clear) | cmd.approx_set;
    const bool tiled = (flags & VRT_CUDA_LIST_MASK) == VRT_CUDA_LIST_REFERENCE || (flags & VRT_CUDA_LIST_MASK) == VRT_CUDA_LIST_REFERENCE_BOUND;

    float angle = cmd.initial_rot; // accumulated rotation about +y (main.cpp:252-255, 330-334)
    double total_time = 0.0;
    for (uint64_t frame = 1; frame <= cmd.nr_frames; ++frame)
    {
        vrt_cuda_frame f;
        std::memset(&f, 0, sizeof(f));
        vrt_host_app_camera(cmd.camera_offset, cmd.focal_length, angle, f.view, f.origin);
        f.width = width; f.height = height;
        f.tiles_x = f.tiles_y = tiled ? (uint32_t)cmd.tiles : 1u;
        f.flags = flags;
        f.bound_sigmas = cmd.bound;

        const double t0 = now_ms();
        vrt_cuda_stats st;
        std::memset(&st, 0, sizeof(st));
        double tiling_time = 0.0, draw_time = 0.0;
        if (n_gpus == 1)
        {
            if (vrt_cuda_tile(ctx[0], &f) != VRT_CUDA_OK) DIE(ctx[0], "vrt_cuda_tile");
            const double t1 = now_ms();
            if (vrt_cuda_render(ctx[0], &f, image.data(), nullptr, &st) != VRT_CUDA_OK) DIE(ctx[0], "vrt_cuda_render");
            tiling_time = t1 - t0;
            draw_time = now_ms() - t1;
        }
        else
        {
            // row bands balanced by the per-row cost of the full-frame lists (built once on GPU 0)
            if (vrt_cuda_tile(ctx[0], &f) != VRT_CUDA_OK) DIE(ctx[0], "vrt_cuda_tile");
            uint32_t n_rows = 0, row_px = 0;
            vrt_cuda_row_costs(ctx[0], nullptr, 0, &n_rows, &row_px);
            std::vector<double> cost(n_rows);
            if (vrt_cuda_row_costs(ctx[0], cost.data(), n_rows, nullptr, nullptr) != VRT_CUDA_OK) DIE(ctx[0], "vrt_cuda_row_costs");
            const uint32_t align = tiled ? height / (uint32_t)cmd.tiles : 16u; // band edges on reference-tile rows
            const uint32_t group = align / row_px > 0 ? align / row_px : 1;
            std::vector<double> gcost((n_rows + group - 1) / group, 0.0);
            double sum = 0.0;
            for (uint32_t i = 0; i < n_rows; ++i) { gcost[i / group] += cost[i]; sum += cost[i]; }
            for (double &c : gcost) c += (sum > 0 ? sum : 1.0) * 1e-6;
            std::vector<uint32_t> bounds(n_gpus + 1);
            vrt_host_row_bands(gcost.data(), (uint32_t)gcost.size(), (uint32_t)n_gpus, bounds.data());
            // one slice size on every GPU (the full frame's choice at a 1/n share): the bands then compose to exactly the
            // image a single GPU renders
            int slice = 0;
            if (vrt_cuda_auto_slice(ctx[0], 1.0 / n_gpus, &slice) != VRT_CUDA_OK) DIE(ctx[0], "vrt_cuda_auto_slice");
            for (int d = 0; d < n_gpus; ++d) vrt_cuda_set_slice(ctx[d], slice);
            const double t1 = now_ms();
            std::vector<std::thread> workers;
            std::vector<vrt_cuda_stats> sts(n_gpus);
            for (int d = 0; d < n_gpus; ++d)
                workers.emplace_back([&, d]() {
                    vrt_cuda_frame fb = f;
                    fb.row_begin = std::min(bounds[d] * group * row_px, height);
                    fb.row_end = d + 1 == n_gpus ? height : std::min(bounds[d + 1] * group * row_px, height);
                    if (fb.row_begin >= fb.row_end) return;
                    if (vrt_cuda_frame_render(ctx[d], &fb, image.data(), nullptr, &sts[d]) != VRT_CUDA_OK) DIE(ctx[d], "vrt_cuda_frame_render");
                });
            for (std::thread &w : workers) w.join();
            for (const vrt_cuda_stats &s : sts) { st.terms_executed += s.terms_executed; st.terms_listed += s.terms_listed; st.list_entries += s.list_entries; }
            tiling_time = t1 - t0;
            draw_time = now_ms() - t1;
        }

        if (cmd.outfile != nullptr)
        {
            // name_<frame>.ext for multi-frame runs (main.cpp:299-305)
            const std::string out(cmd.outfile);
            const size_t dot = out.find_last_of('.');
            const std::string stem = dot == std::string::npos ? out : out.substr(0, dot), ext = dot == std::string::npos ? "png" : out.substr(dot + 1);
            const std::string name = cmd.nr_frames > 1 ? stem + "_" + std::to_string(frame) + "." + ext : stem + "." + ext;
            if (vrt_host_write_png(name.c_str(), width, height, image.data()) != 0) std::fprintf(stderr, "[ ERROR ]\tcannot write %s\n", name.c_str());
        }
        if (cmd.nr_frames == 1)
        {
            std::printf("TIME: %g ms\n", draw_time + tiling_time);
            std::printf("[ INFO ]\t%llu Gaussians, %.4g evaluations listed, %.4g executed, %.4g evaluations/s\n", (unsigned long long)gaussians.size(),
                        st.terms_listed, st.terms_executed, st.terms_executed / ((draw_time + tiling_time) * 1e-3));
        }
        total_time += draw_time + tiling_time;
        if (cmd.nr_frames == frame && cmd.nr_frames > 1) std::printf("AVG. TIME: %g ms (%llu frames)\n", total_time / cmd.nr_frames, (unsigned long long)cmd.nr_frames);
        angle += cmd.rot / cmd.nr_frames; // main.cpp:330
    }
    if (pinned) vrt_cuda_unpin_buffer(ctx[0], image.data());
    for (vrt_cuda_ctx *c : ctx) vrt_cuda_destroy(c);
    return EXIT_SUCCESS;
}
