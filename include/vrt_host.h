/* vrt_host.h -- host-side (CPU only, no CUDA) helpers of the B200 `vrt` render path.
 *
 * These restate the *callers' side* of the reference's hot path -- scene construction, the camera
 * and the framebuffer writer of src/volumetric-ray-tracer/main.cpp -- behind a plain C ABI so the
 * headless `volumetric-ray-tracer` (C++), the Python tests and bench.py all share one
 * implementation.  Nothing here renders; rendering is include/vrt_cuda.h.
 *
 * All citations are relative to the reference repository root.
 */
#ifndef VRT_HOST_H
#define VRT_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One Gaussian = 10 packed floats, the memory layout of vrt::gaussian_t (src/vrt/types.h:195-200):
 * albedo.xyzw, mu.xyzw, sigma, magnitude.  `std::vector<gaussian_t>::data()` can be passed as is. */
#define VRT_GAUSSIAN_FLOATS 10

/* Synthetic `-g <dim>` grid of src/volumetric-ray-tracer/main.cpp:196-204 (dim*dim Gaussians at z = 1,
 * sigma = 1/(2 dim), magnitude 1, albedo (1-t, 0, t, 1), t = (i*dim+j)/dim^2).  Loop indices are u8 in
 * the reference, so dim must be <= 255.  Returns the Gaussian count or 0 on bad arguments. */
uint64_t vrt_host_scene_grid(uint32_t dim, float *aos_out);

/* Same grid with explicit sigma / magnitude (scene of tests/img-error.cpp:18-26: dim 16, sigma 1/4,
 * magnitude 3). */
uint64_t vrt_host_scene_grid_ex(uint32_t dim, float sigma, float magnitude, float *aos_out);

/* The three hand-written Gaussians of tests/transmittance.cpp:9.  aos_out: 3 x 10 floats. */
uint64_t vrt_host_scene_transmittance_test(float *aos_out);

/* Frustum-filling random scene for BASELINE configs 4/5; not part of the reference
 * CLI.  Counter-based RNG: u(index, field) = top 24 bits of splitmix64(seed * 2^40 + index * 8 + field).
 * z ~ U(0,2); (x, y) = (U(-1,1), U(-1,1)) * (z + 4); sigma = 10^U(log10_sigma_lo, log10_sigma_hi);
 * optical depth through the centre tau ~ U(0.2, 1.5), magnitude = tau / (sigma sqrt(2 pi));
 * albedo rgb ~ U(0,1), albedo.w = 1.  Returns n. */
uint64_t vrt_host_scene_synthetic(uint64_t n, uint64_t seed, float log10_sigma_lo, float log10_sigma_hi, float *aos_out);

/* OBJ vertices -> Gaussians, read_from_obj of src/vrt/gaussians-from-file.cpp:7-44 (sigma by vertex count,
 * albedo = 0.5 * normalize(p) + 0.5, albedo.w = 1, magnitude 1).  Returns the vertex count (copies at most
 * `cap` Gaussians; call with cap = 0 to size the buffer) or UINT64_MAX if the file cannot be read. */
uint64_t vrt_host_read_obj(const char *path, float *aos_out, uint64_t cap);

/* camera_t::turn + camera_t::update of src/vrt/camera.cpp:7-23, :52: column-major view matrix
 * translate(lookAt_RH(pos, pos + front, up), focal * front) for yaw / pitch in degrees. */
void vrt_host_view_matrix(const float pos[3], float yaw_deg, float pitch_deg, float focal, float view16_out[16]);

/* The app's orbiting camera, src/volumetric-ray-tracer/main.cpp:248-255 and :330-334: position (0, 0,
 * camera_offset) rotated by `rotation_deg` about +y, yaw = -90 - rotation_deg.  origin4_out.w = 0. */
void vrt_host_app_camera(float camera_offset, float focal, float rotation_deg, float view16_out[16], float origin4_out[4]);

/* Work-balanced split of `n_rows` tile rows into `n_parts` contiguous bands:
 * bounds_out has n_parts + 1 entries, bounds_out[0] = 0, bounds_out[n_parts] = n_rows, chosen so that the
 * largest band cost (sum of row_cost) is minimal among prefix-sum splits.  Returns 0, or -1 on bad input. */
int vrt_host_row_bands(const double *row_cost, uint32_t n_rows, uint32_t n_parts, uint32_t *bounds_out);

/* Writes the packed 0xAARRGGBB framebuffer as an 8-bit RGBA PNG exactly the way the reference hands it to
 * stbi_write_png (src/volumetric-ray-tracer/main.cpp:306): the little-endian bytes of each u32 are taken
 * as R,G,B,A, i.e. PNG red = B channel.  Stored (uncompressed) deflate blocks.  Returns 0 on success. */
int vrt_host_write_png(const char *path, uint32_t width, uint32_t height, const uint32_t *image);

#ifdef __cplusplus
}
#endif
#endif /* VRT_HOST_H */
