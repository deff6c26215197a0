// vrt_types.hpp -- stand-alone mirrors of the reference's data carriers for the hot path.
//
// Same names, same field order and same memory layout as src/vrt/types.h / src/vrt/camera.h of the reference, written
// from scratch without T-SIMD, GLM or ImGui, so that a host program can be built against include/vrt_cuda.hpp with or
// without the reference tree.  Only what the render entries read is kept:
//     vec4f_t      types.h:19-21      x, y, z, w (w defaults to 0)
//     gaussian_t   types.h:195-200    albedo, mu, sigma, magnitude -- 40 bytes, the C ABI's record layout
//     gaussians_t  types.h:266-270    the AoS vector (the SoA mirror `soa_gaussians` is a CPU-SIMD detail and is omitted)
//     tiles_t      types.h:272-287    per-tile gaussians_t, tw, th, w = ceil(2/tw), h = ceil(2/th)
//     camera_t     camera.h:20-44     position, front, up, right, view_matrix, focal_length, w, h; no projection-plane arrays:
//                                     the GPU regenerates the plane point of a pixel from inverse(view_matrix)
#pragma once

#include <cmath>
#include <cstdint>
#include <vector>

#include "vrt_host.h"

namespace vrt
{
struct vec4f_t
{
    float x, y, z, w = 0.f;
};

struct gaussian_t
{
    vec4f_t albedo;
    vec4f_t mu;
    float sigma;
    float magnitude;
};
static_assert(sizeof(gaussian_t) == 40, "gaussian_t is the 40-byte record of the C ABI");

struct gaussians_t
{
    std::vector<gaussian_t> gaussians;
};

struct tiles_t
{
    std::vector<gaussians_t> gaussians;
    float tw, th;
    uint64_t w, h;
    tiles_t(const std::vector<gaussians_t> &g, float tw_, float th_)
        : gaussians(g), tw(tw_), th(th_), w((uint64_t)std::ceil(2.f / tw_)), h((uint64_t)std::ceil(2.f / th_)) {}
};

/// Column-major 4x4 with m[col][row] access, the only part of glm::mat4 the render path touches.
struct mat4_t
{
    float c[4][4];
    float *operator[](int col) { return c[col]; }
    const float *operator[](int col) const { return c[col]; }
};

struct camera_t
{
    float position[3];
    float yaw, pitch;
    mat4_t view_matrix;
    float focal_length;
    uint64_t w, h;

    camera_t(const float pos[3], float yaw_ = -90.f, float pitch_ = 0.f, uint64_t width = 256, uint64_t height = 256, float focal = 1.f)
        : yaw(yaw_), pitch(pitch_), focal_length(focal), w(width), h(height)
    {
        position[0] = pos[0]; position[1] = pos[1]; position[2] = pos[2];
        update();
    }
    /// camera_t::turn (camera.cpp:7-23): new yaw / pitch in degrees, then update().
    void turn(float yaw_, float pitch_)
    {
        yaw = yaw_;
        pitch = pitch_;
        update();
    }
    /// camera_t::update (camera.cpp:50-52): view = translate(lookAt(pos, pos + front, up), focal * front).
    void update() { vrt_host_view_matrix(position, yaw, pitch, focal_length, &view_matrix.c[0][0]); }
};
} // namespace vrt
