// vrt_host.cpp -- CPU-side helpers of the B200 `vrt` render path (see include/vrt_host.h).
//
// Restates the caller's side of the reference hot path: scene construction
// (src/volumetric-ray-tracer/main.cpp:196-204, src/vrt/gaussians-from-file.cpp:7-44), the camera
// (src/vrt/camera.cpp:7-23, :52; main.cpp:248-255) and the PNG hand-off (main.cpp:299-307).
// No CUDA, no rendering, no dependency on anything under oracle/.
#include "vrt_host.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

namespace
{
    enum { AR = 0, AG, AB, AW, MX, MY, MZ, MW, SIGMA, MAG };

    struct v3 { float x, y, z; };
    inline v3 cross(v3 a, v3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
    inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
    inline v3 normalize(v3 a) { const float i = 1.f / std::sqrt(dot(a, a)); return {a.x * i, a.y * i, a.z * i}; }

    inline uint64_t splitmix64(uint64_t x)
    {
        x += 0x9E3779B97F4A7C15ull;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        return x ^ (x >> 31);
    }
    // uniform in [0,1) with 24 random bits (exactly representable in fp32)
    inline double u01(uint64_t seed, uint64_t index, uint64_t field)
    {
        return (double)(splitmix64((seed << 40) + index * 8 + field) >> 40) * (1.0 / 16777216.0);
    }

    uint64_t grid(uint32_t dim, float sigma, float magnitude, float *out)
    {
        if (dim == 0 || dim > 255 || out == nullptr) return 0;
        const int d = (int)dim;
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j)
            {
                float *g = out + (size_t)(i * d + j) * VRT_GAUSSIAN_FLOATS;
                const float t = (i * d + j) / (float)(d * d);
                g[AR] = 1.f - t; g[AG] = 0.f; g[AB] = 0.f + t; g[AW] = 1.f;
                g[MX] = -1.f + 1.f / d + i * 1.f / (d / 2.f);
                g[MY] = -1.f + 1.f / d + j * 1.f / (d / 2.f);
                g[MZ] = 1.f; g[MW] = 0.f;
                g[SIGMA] = sigma; g[MAG] = magnitude;
            }
        return (uint64_t)d * d;
    }

    // ---- PNG (stored deflate) ----
    struct CrcTable
    {
        uint32_t t[256];
        CrcTable()
        {
            for (uint32_t i = 0; i < 256; ++i)
            {
                uint32_t c = i;
                for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
                t[i] = c;
            }
        }
    };
    uint32_t crc32(uint32_t crc, const uint8_t *p, size_t n)
    {
        static const CrcTable table; // thread-safe one-time initialisation (the app writes frames from one thread, tests may not)
        crc = ~crc;
        for (size_t i = 0; i < n; ++i) crc = table.t[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
        return ~crc;
    }
    void put32(std::vector<uint8_t> &v, uint32_t x)
    {
        v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x);
    }
    void chunk(std::vector<uint8_t> &png, const char *type, const std::vector<uint8_t> &data)
    {
        put32(png, (uint32_t)data.size());
        const size_t start = png.size();
        png.insert(png.end(), type, type + 4);
        png.insert(png.end(), data.begin(), data.end());
        put32(png, crc32(0, png.data() + start, png.size() - start));
    }
}

extern "C"
{
    uint64_t vrt_host_scene_grid(uint32_t dim, float *aos_out)
    {
        return grid(dim, dim ? 1.f / (2 * (int)dim) : 0.f, 1.f, aos_out);
    }

    uint64_t vrt_host_scene_grid_ex(uint32_t dim, float sigma, float magnitude, float *aos_out)
    {
        return grid(dim, sigma, magnitude, aos_out);
    }

    uint64_t vrt_host_scene_transmittance_test(float *aos_out)
    {
        const float g[3][VRT_GAUSSIAN_FLOATS] = {
            {0.f, 1.f, 0.f, .1f, .3f, .3f, .5f, 0.f, 0.1f, 2.f},
            {0.f, 0.f, 1.f, .7f, -.3f, -.3f, 0.f, 0.f, 0.4f, .7f},
            {1.f, 0.f, 0.f, 1.f, 0.f, 0.f, 2.f, 0.f, .75f, 1.f}};
        std::memcpy(aos_out, g, sizeof(g));
        return 3;
    }

    uint64_t vrt_host_scene_synthetic(uint64_t n, uint64_t seed, float log10_sigma_lo, float log10_sigma_hi, float *aos_out)
    {
        if (aos_out == nullptr) return 0;
        for (uint64_t i = 0; i < n; ++i)
        {
            float *g = aos_out + i * VRT_GAUSSIAN_FLOATS;
            const double z = 2.0 * u01(seed, i, 0);
            const double u = 2.0 * u01(seed, i, 1) - 1.0, v = 2.0 * u01(seed, i, 2) - 1.0;
            const double ls = (double)log10_sigma_lo + ((double)log10_sigma_hi - (double)log10_sigma_lo) * u01(seed, i, 3);
            // 10^ls through exp2 of an fp32-rounded exponent keeps the result reproducible across libms
            const float sigma = (float)std::exp2((double)(float)(ls * 3.321928094887362));
            const double tau = 0.2 + 1.3 * u01(seed, i, 4);
            g[AR] = (float)u01(seed, i, 5); g[AG] = (float)u01(seed, i, 6); g[AB] = (float)u01(seed, i, 7); g[AW] = 1.f;
            g[MX] = (float)(u * (z + 4.0)); g[MY] = (float)(v * (z + 4.0)); g[MZ] = (float)z; g[MW] = 0.f;
            g[SIGMA] = sigma;
            g[MAG] = (float)(tau / ((double)sigma * 2.5066282746310002));
        }
        return n;
    }

    uint64_t vrt_host_read_obj(const char *path, float *aos_out, uint64_t cap)
    {
        FILE *f = std::fopen(path, "r");
        if (!f) return std::numeric_limits<uint64_t>::max();
        char line[4096];
        uint64_t n = 0;
        while (std::fgets(line, sizeof(line), f))
        {
            float x, y, z;
            if (line[0] != 'v' || (line[1] != ' ' && line[1] != '\t')) continue;
            if (std::sscanf(line + 2, "%f %f %f", &x, &y, &z) != 3) continue;
            if (n < cap && aos_out != nullptr)
            {
                float *g = aos_out + n * VRT_GAUSSIAN_FLOATS;
                const float norm = std::sqrt(x * x + y * y + z * z);
                g[AR] = x / norm * 0.5f + 0.5f; g[AG] = y / norm * 0.5f + 0.5f; g[AB] = z / norm * 0.5f + 0.5f; g[AW] = 1.0f;
                g[MX] = x; g[MY] = y; g[MZ] = z; g[MW] = 0.f;
                g[SIGMA] = 0.f; g[MAG] = 1.0f;
            }
            ++n;
        }
        std::fclose(f);
        const float sig = n < 300 ? 0.3f : (n < 1000 ? 0.15f : 0.05f);
        for (uint64_t i = 0; i < n && i < cap && aos_out != nullptr; ++i) aos_out[i * VRT_GAUSSIAN_FLOATS + SIGMA] = sig;
        return n;
    }

    void vrt_host_view_matrix(const float pos[3], float yaw_deg, float pitch_deg, float focal, float view16_out[16])
    {
        const float d2r = 0.017453292519943295f;
        if (pitch_deg > 89.f) pitch_deg = 89.f;
        if (pitch_deg < -89.f) pitch_deg = -89.f;
        const v3 eye{pos[0], pos[1], pos[2]};
        const v3 front = normalize({std::cos(yaw_deg * d2r) * std::cos(pitch_deg * d2r), std::sin(pitch_deg * d2r),
                                    std::sin(yaw_deg * d2r) * std::cos(pitch_deg * d2r)});
        const v3 right = normalize(cross(front, {0.f, 1.f, 0.f}));
        const v3 up = normalize(cross(right, front));
        // right-handed look-at towards eye + front
        const v3 f = normalize({(eye.x + front.x) - eye.x, (eye.y + front.y) - eye.y, (eye.z + front.z) - eye.z});
        const v3 s = normalize(cross(f, up));
        const v3 u = cross(s, f);
        float m[16] = {s.x, u.x, -f.x, 0.f, s.y, u.y, -f.y, 0.f, s.z, u.z, -f.z, 0.f, -dot(s, eye), -dot(u, eye), dot(f, eye), 1.f};
        // post-multiplied translation by focal * front
        const v3 t{focal * front.x, focal * front.y, focal * front.z};
        for (int i = 0; i < 4; ++i) m[12 + i] = m[i] * t.x + m[4 + i] * t.y + m[8 + i] * t.z + m[12 + i];
        std::memcpy(view16_out, m, sizeof(m));
    }

    void vrt_host_app_camera(float camera_offset, float focal, float rotation_deg, float view16_out[16], float origin4_out[4])
    {
        const float a = rotation_deg * 0.017453292519943295f;
        const float pos[3] = {std::sin(a) * camera_offset, 0.f, std::cos(a) * camera_offset};
        vrt_host_view_matrix(pos, -90.f - rotation_deg, 0.f, focal, view16_out);
        origin4_out[0] = pos[0]; origin4_out[1] = pos[1]; origin4_out[2] = pos[2]; origin4_out[3] = 0.f;
    }

    int vrt_host_row_bands(const double *row_cost, uint32_t n_rows, uint32_t n_parts, uint32_t *bounds_out)
    {
        if (!row_cost || !bounds_out || n_parts == 0) return -1;
        std::vector<double> prefix(n_rows + 1, 0.0);
        for (uint32_t i = 0; i < n_rows; ++i)
        {
            if (!(row_cost[i] >= 0.0)) return -1;
            prefix[i + 1] = prefix[i] + row_cost[i];
        }
        // best[p][i] = minimal achievable max band cost when the first i rows form p bands
        const double inf = std::numeric_limits<double>::infinity();
        std::vector<std::vector<double>> best(n_parts + 1, std::vector<double>(n_rows + 1, inf));
        std::vector<std::vector<uint32_t>> cut(n_parts + 1, std::vector<uint32_t>(n_rows + 1, 0));
        best[0][0] = 0.0;
        for (uint32_t p = 1; p <= n_parts; ++p)
            for (uint32_t i = 0; i <= n_rows; ++i)
                for (uint32_t k = 0; k <= i; ++k)
                {
                    if (best[p - 1][k] == inf) continue;
                    const double c = std::fmax(best[p - 1][k], prefix[i] - prefix[k]);
                    // ties go to the later cut so leading bands are never needlessly empty
                    if (c < best[p][i] || (c == best[p][i] && k > cut[p][i])) { best[p][i] = c; cut[p][i] = k; }
                }
        uint32_t i = n_rows;
        bounds_out[n_parts] = n_rows;
        for (uint32_t p = n_parts; p >= 1; --p)
        {
            i = cut[p][i];
            bounds_out[p - 1] = i;
        }
        return 0;
    }

    int vrt_host_write_png(const char *path, uint32_t width, uint32_t height, const uint32_t *image)
    {
        if (!path || !image || width == 0 || height == 0) return -1;
        // raw scanlines: filter byte 0 + the u32 pixels' little-endian bytes as R,G,B,A (main.cpp:306)
        const size_t stride = (size_t)width * 4 + 1;
        std::vector<uint8_t> raw(stride * height);
        for (uint32_t y = 0; y < height; ++y)
        {
            raw[y * stride] = 0;
            std::memcpy(&raw[y * stride + 1], image + (size_t)y * width, (size_t)width * 4);
        }
        std::vector<uint8_t> z;
        z.push_back(0x78); z.push_back(0x01);
        uint32_t a = 1, b = 0;
        for (size_t off = 0; off < raw.size();)
        {
            const size_t len = std::min<size_t>(65535, raw.size() - off);
            z.push_back(off + len == raw.size() ? 1 : 0);
            z.push_back(len & 0xFF); z.push_back(len >> 8);
            z.push_back(~len & 0xFF); z.push_back((~len >> 8) & 0xFF);
            z.insert(z.end(), raw.begin() + off, raw.begin() + off + len);
            for (size_t i = 0; i < len; ++i) { a = (a + raw[off + i]) % 65521u; b = (b + a) % 65521u; }
            off += len;
        }
        put32(z, (b << 16) | a);
        std::vector<uint8_t> png = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
        std::vector<uint8_t> ihdr;
        put32(ihdr, width); put32(ihdr, height);
        ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
        chunk(png, "IHDR", ihdr);
        chunk(png, "IDAT", z);
        chunk(png, "IEND", {});
        FILE *f = std::fopen(path, "wb");
        if (!f) return -2;
        const size_t w = std::fwrite(png.data(), 1, png.size(), f);
        std::fclose(f);
        return w == png.size() ? 0 : -3;
    }
}
