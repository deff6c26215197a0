/* oracle/vrt_oracle.c -- CPU restatement of the reference's `vrt` hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load libvrt_oracle.so; the product (libvrt_cuda.so) never does and has
 * no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below against the unmodified
 * reference compiled in place (oracle/_ref, see oracle/Makefile) and against golden vectors that
 * were generated from that reference build (tests/golden/, generator tests/golden/make_golden.py).
 *
 * All citations are relative to /root/reference/.  Written from the algorithm, not copied:
 * plain C, scalar loops, no SIMD, no templates.  Compiled with -fno-fast-math -ffp-contract=off
 * so the fp32 entry points are a well-defined IEEE evaluation of the reference's formulas (the
 * reference itself is built -ffast-math, so it is only reproducible to rounding noise; see
 * DESIGN.md "numerical conditioning").  The *_f64 entry points evaluate the same formulas in
 * double and serve as the arbiter when two fp32 evaluations disagree.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* gaussian_t: src/vrt/types.h:195-200 -- albedo xyzw, mu xyzw, sigma, magnitude = 10 floats. */
enum { G_AR = 0, G_AG, G_AB, G_AW, G_MX, G_MY, G_MZ, G_MW, G_SIGMA, G_MAG, G_STRIDE };

/* src/vrt/rt.h:18-20 */
static const float SQRT_2_PI_F = 0.7978845608028654f;
#define INV_SQRT_2_PI_F (1.f / SQRT_2_PI_F)
static const float SQRT_2_F = 1.41421356237309504880f;

/* ---------------------------------------------------------------- erf variants -------------- */

/* src/vrt/approx.cpp:90-99 -- Abramowitz & Stegun 7.1.27 with the reference's coefficients. */
ORC_API float orc_as_erf_f32(float x)
{
    const float sign = (float)((x >= 0) - (x < 0));
    x *= sign;
    const float denom = (((0.078108f * x + 0.000972f) * x + 0.230389f) * x + 0.278393f) * x + 1.f;
    const float denom2 = denom * denom;
    const float val = 1.f - 1.f / (denom2 * denom2);
    return val * sign;
}

static double as_erf_f64(double x)
{
    const double sign = (double)((x >= 0) - (x < 0));
    x *= sign;
    /* the coefficients are the reference's fp32 literals, widened */
    const double denom = ((((double)0.078108f * x + (double)0.000972f) * x + (double)0.230389f) * x + (double)0.278393f) * x + 1.0;
    const double denom2 = denom * denom;
    return (1.0 - 1.0 / (denom2 * denom2)) * sign;
}

/* ---------------------------------------------------------------- alternative approximations */
#include "../include/vrt_approx_tables.h"

/* Piecewise-cubic lookup shared by spline_erf (src/vrt/approx.cpp:9-23) and spline_exp (:141-163):
 * x <= first knot -> `below`, x >= last knot -> `above`, else the segment whose half-open interval holds x. */
static float spline_f32(const float *knot, const float (*coef)[4], int nseg, float below, float above, float x)
{
    if (x <= knot[0]) return below;
    for (int i = 0; i < nseg; ++i)
        if (x < knot[i + 1])
        {
            const float d = x - knot[i];
            return ((coef[i][0] * d + coef[i][1]) * d + coef[i][2]) * d + coef[i][3];
        }
    return above;
}
static double spline_f64(const float *knot, const float (*coef)[4], int nseg, double below, double above, double x)
{
    if (x <= (double)knot[0]) return below;
    for (int i = 0; i < nseg; ++i)
        if (x < (double)knot[i + 1])
        {
            const double d = x - (double)knot[i];
            return (((double)coef[i][0] * d + (double)coef[i][1]) * d + (double)coef[i][2]) * d + (double)coef[i][3];
        }
    return above;
}

ORC_API float orc_spline_erf_f32(float x) { return spline_f32(vrt_spline_erf_knot, vrt_spline_erf_coef, VRT_SPLINE_ERF_SEGMENTS, -1.f, 1.f, x); }
ORC_API float orc_spline_exp_f32(float x) { return spline_f32(vrt_spline_exp_knot, vrt_spline_exp_coef, VRT_SPLINE_EXP_SEGMENTS, 0.f, 1.f, x); }

/* src/vrt/approx.cpp:45-56: the negative half of the erf spline (segments 0..3, then segment 4 for everything up to 0)
 * evaluated at -|x| and mirrored; SIGN(0) = +1 there (approx.cpp:5). */
ORC_API float orc_spline_erf_mirror_f32(float x)
{
    const float sign = (float)((x >= 0) - (x < 0));
    const float m = -sign * x; /* -|x| */
    if (m <= vrt_spline_erf_knot[0]) return sign;
    int i = 0;
    while (i < 4 && !(m < vrt_spline_erf_knot[i + 1])) ++i;
    const float d = m - vrt_spline_erf_knot[i];
    const float v = ((vrt_spline_erf_coef[i][0] * d + vrt_spline_erf_coef[i][1]) * d + vrt_spline_erf_coef[i][2]) * d + vrt_spline_erf_coef[i][3];
    return -sign * v;
}
static double spline_erf_mirror_f64(double x)
{
    const double sign = (double)((x >= 0) - (x < 0));
    const double m = -sign * x;
    if (m <= (double)vrt_spline_erf_knot[0]) return sign;
    int i = 0;
    while (i < 4 && !(m < (double)vrt_spline_erf_knot[i + 1])) ++i;
    const double d = m - (double)vrt_spline_erf_knot[i];
    const double v = (((double)vrt_spline_erf_coef[i][0] * d + (double)vrt_spline_erf_coef[i][1]) * d + (double)vrt_spline_erf_coef[i][2]) * d + (double)vrt_spline_erf_coef[i][3];
    return -sign * v;
}

/* src/vrt/approx.cpp:68-77: ten terms of the Maclaurin series of erf, coefficients (-1)^n / (n! (2n+1)) as fp32,
 * saturated to +-1 outside (-2, 2). */
static const double taylor_den[10] = {1.0, -3.0, 10.0, -42.0, 216.0, -1320.0, 9360.0, -75600.0, 685440.0, -6894720.0};
static const float TWO_INV_SQRTPI_F = 2.f * 0.5641895835477563f;
ORC_API float orc_taylor_erf_f32(float x)
{
    if (x <= -2.f) return -1.f;
    if (x >= 2.f) return 1.f;
    float p = (float)(1.0 / taylor_den[9]);
    for (int n = 8; n >= 0; --n) p = p * x * x + (float)(1.0 / taylor_den[n]);
    return TWO_INV_SQRTPI_F * p * x;
}
static double taylor_erf_f64(double x)
{
    if (x <= -2.0) return -1.0;
    if (x >= 2.0) return 1.0;
    double p = (double)(float)(1.0 / taylor_den[9]);
    for (int n = 8; n >= 0; --n) p = p * x * x + (double)(float)(1.0 / taylor_den[n]);
    return (double)TWO_INV_SQRTPI_F * p * x;
}

/* src/vrt/approx.cpp:112-127, Schraudolph's exp: the float a x + b, converted to an integer, IS the bit pattern of the
 * result.  The range clamp is the one the reference compiles in unless NDEBUG is defined (its CMake build does not define
 * it); outside [2^23, 255 * 2^23] the unclamped conversion is undefined behaviour, so the clamp is the only definition.
 * The scalar function truncates, the SIMD one (:130-137, cvts) rounds to nearest: `nearest` selects. */
static float fast_exp_impl(float x, int nearest)
{
    const float a = (float)(1 << 23) / 0.6931471805599453f;
    const float b = (float)(1 << 23) * (127.f - 0.043677448f);
    const float lo = (float)(1 << 23), hi = (float)(1 << 23) * 255.f;
    float y = a * x + b;
    if (y < lo || y > hi) y = (y < lo) ? 0.f : hi;
    const uint32_t n = nearest ? (uint32_t)llrintf(y) : (uint32_t)y;
    float r;
    memcpy(&r, &n, sizeof(r));
    return r;
}
ORC_API float orc_fast_exp_f32(float x) { return fast_exp_impl(x, 0); }
ORC_API float orc_fast_exp_simd_f32(float x) { return fast_exp_impl(x, 1); }

/* Variant code = erf id | exp id << 4.
 * erf id 0: libm erff (template default, src/vrt/rt.h:32)   1: Abramowitz-Stegun   2: spline   3: mirrored spline   4: Taylor
 * exp id 0: libm expf                                      1: fast_exp (SIMD rounding)        2: spline_exp
 * Exp applies to c_bar and to the final exp(T) (rt.h:45, 53 / :115, 126); the density G_q(x) always uses libm exp
 * (types.h:204-208; the SIMD pdf's template default, rt.h:218). */
static inline float erf_variant_f32(int variant, float x)
{
    switch (variant & 15)
    {
    case 1: return orc_as_erf_f32(x);
    case 2: return orc_spline_erf_f32(x);
    case 3: return orc_spline_erf_mirror_f32(x);
    case 4: return orc_taylor_erf_f32(x);
    default: return erff(x);
    }
}
static inline double erf_variant_f64(int variant, double x)
{
    switch (variant & 15)
    {
    case 1: return as_erf_f64(x);
    case 2: return spline_f64(vrt_spline_erf_knot, vrt_spline_erf_coef, VRT_SPLINE_ERF_SEGMENTS, -1.0, 1.0, x);
    case 3: return spline_erf_mirror_f64(x);
    case 4: return taylor_erf_f64(x);
    default: return erf(x);
    }
}
static inline float exp_variant_f32(int variant, float x)
{
    switch ((variant >> 4) & 15)
    {
    case 1: return orc_fast_exp_simd_f32(x);
    case 2: return orc_spline_exp_f32(x);
    default: return expf(x);
    }
}
static inline double exp_variant_f64(int variant, double x)
{
    switch ((variant >> 4) & 15)
    {
    case 1: return (double)orc_fast_exp_simd_f32((float)x); /* the bit trick is an fp32 construction */
    case 2: return spline_f64(vrt_spline_exp_knot, vrt_spline_exp_coef, VRT_SPLINE_EXP_SEGMENTS, 0.0, 1.0, x);
    default: return exp(x);
    }
}

/* tabulation entry for tests (the functions tests/accuracy.cpp tabulates): fn 0 spline_erf, 1 spline_erf_mirror,
 * 2 taylor_erf, 3 abramowitz_stegun_erf, 4 erff, 5 expf, 6 fast_exp (scalar, truncating), 7 spline_exp, 8 fast_exp (SIMD rounding) */
ORC_API void orc_approx_table(int fn, const float *x, uint64_t n, float *y)
{
    for (uint64_t i = 0; i < n; ++i)
    {
        const float v = x[i];
        switch (fn)
        {
        case 0: y[i] = orc_spline_erf_f32(v); break;
        case 1: y[i] = orc_spline_erf_mirror_f32(v); break;
        case 2: y[i] = orc_taylor_erf_f32(v); break;
        case 3: y[i] = orc_as_erf_f32(v); break;
        case 4: y[i] = erff(v); break;
        case 5: y[i] = expf(v); break;
        case 6: y[i] = orc_fast_exp_f32(v); break;
        case 7: y[i] = orc_spline_exp_f32(v); break;
        default: y[i] = orc_fast_exp_simd_f32(v); break;
        }
    }
}

/* ---------------------------------------------------------------- transmittance ------------- */

/* src/vrt/rt.h:32-54.  o, n: 4 floats each (w participates in the dot products, types.h:56-59). */
ORC_API float orc_transmittance_f32(const float *o, const float *n, float s, const float *g, uint64_t count, int variant)
{
    float T = 0.f;
    for (uint64_t j = 0; j < count; ++j)
    {
        const float *q = g + j * G_STRIDE;
        const float cx = q[G_MX] - o[0], cy = q[G_MY] - o[1], cz = q[G_MZ] - o[2], cw = q[G_MW] - o[3];
        const float mu_bar = cx * n[0] + cy * n[1] + cz * n[2] + cw * n[3];
        const float oc_sqnorm = cx * cx + cy * cy + cz * cz + cw * cw;
        const float mb2 = mu_bar * mu_bar;
        const float sigma = q[G_SIGMA];
        const float inv_2_sigma2 = 1.f / (2.f * sigma * sigma);
        const float c_bar = q[G_MAG] * exp_variant_f32(variant, -((oc_sqnorm - mb2) * inv_2_sigma2));
        const float sqrt_2_sig = SQRT_2_F * sigma;
        const float mu_bar_n = mu_bar / sqrt_2_sig;
        const float s_n = s / sqrt_2_sig;
        const float erf1 = erf_variant_f32(variant, -mu_bar_n);
        const float erf2 = erf_variant_f32(variant, s_n - mu_bar_n);
        T += sigma * c_bar * INV_SQRT_2_PI_F * (erf1 - erf2);
    }
    return exp_variant_f32(variant, T);
}

static double transmittance_f64_d(const float *o, const double *n, double s, const float *g, uint64_t count, int variant)
{
    double T = 0.0;
    for (uint64_t j = 0; j < count; ++j)
    {
        const float *q = g + j * G_STRIDE;
        const double cx = (double)q[G_MX] - o[0], cy = (double)q[G_MY] - o[1], cz = (double)q[G_MZ] - o[2], cw = (double)q[G_MW] - o[3];
        const double mu_bar = cx * n[0] + cy * n[1] + cz * n[2] + cw * n[3];
        const double oc_sqnorm = cx * cx + cy * cy + cz * cz + cw * cw;
        const double sigma = q[G_SIGMA];
        const double c_bar = (double)q[G_MAG] * exp_variant_f64(variant, -((oc_sqnorm - mu_bar * mu_bar) / (2.0 * sigma * sigma)));
        const double sqrt_2_sig = (double)SQRT_2_F * sigma;
        const double erf1 = erf_variant_f64(variant, -mu_bar / sqrt_2_sig);
        const double erf2 = erf_variant_f64(variant, s / sqrt_2_sig - mu_bar / sqrt_2_sig);
        T += sigma * c_bar * (double)INV_SQRT_2_PI_F * (erf1 - erf2);
    }
    return exp_variant_f64(variant, T);
}

ORC_API double orc_transmittance_f64(const float *o, const float *n, double s, const float *g, uint64_t count, int variant)
{
    const double nd[4] = {n[0], n[1], n[2], n[3]};
    return transmittance_f64_d(o, nd, s, g, count, variant);
}

/* ---------------------------------------------------------------- radiance ------------------ */

/* gaussian_t::pdf, src/vrt/types.h:204-208 */
static float pdf_f32(const float *q, const float *x)
{
    const float dx = x[0] - q[G_MX], dy = x[1] - q[G_MY], dz = x[2] - q[G_MZ], dw = x[3] - q[G_MW];
    return q[G_MAG] * expf(-(dx * dx + dy * dy + dz * dz + dw * dw) / (2 * q[G_SIGMA] * q[G_SIGMA]));
}

/* src/vrt/rt.h:146-164: 5 samples s = mu_bar_q + k*sigma_q, k = -4..0. out = x,y,z,w. */
ORC_API void orc_radiance_f32(const float *o, const float *n, const float *g, uint64_t count, int variant, float *out)
{
    float L[4] = {0.f, 0.f, 0.f, 0.f};
    for (uint64_t i = 0; i < count; ++i)
    {
        const float *q = g + i * G_STRIDE;
        const float lambda = q[G_SIGMA];
        float inner = 0.f;
        for (int k = -4; k <= 0; ++k)
        {
            const float s = (q[G_MX] - o[0]) * n[0] + (q[G_MY] - o[1]) * n[1] + (q[G_MZ] - o[2]) * n[2] + (q[G_MW] - o[3]) * n[3] + (float)k * lambda;
            const float T = orc_transmittance_f32(o, n, s, g, count, variant);
            const float x[4] = {o[0] + n[0] * s, o[1] + n[1] * s, o[2] + n[2] * s, o[3] + n[3] * s};
            inner += pdf_f32(q, x) * T * lambda;
        }
        for (int c = 0; c < 4; ++c) L[c] = L[c] + q[G_AR + c] * inner;
    }
    memcpy(out, L, sizeof(L));
}

static void radiance_f64_d(const float *o, const double *n, const float *g, uint64_t count, int variant, double *out)
{
    double L[4] = {0, 0, 0, 0};
    for (uint64_t i = 0; i < count; ++i)
    {
        const float *q = g + i * G_STRIDE;
        const double lambda = q[G_SIGMA];
        double inner = 0.0;
        for (int k = -4; k <= 0; ++k)
        {
            const double s = ((double)q[G_MX] - o[0]) * n[0] + ((double)q[G_MY] - o[1]) * n[1] + ((double)q[G_MZ] - o[2]) * n[2] + ((double)q[G_MW] - o[3]) * n[3] + k * lambda;
            const double T = transmittance_f64_d(o, n, s, g, count, variant);
            const double dx = o[0] + n[0] * s - q[G_MX], dy = o[1] + n[1] * s - q[G_MY], dz = o[2] + n[2] * s - q[G_MZ], dw = o[3] + n[3] * s - q[G_MW];
            inner += (double)q[G_MAG] * exp(-(dx * dx + dy * dy + dz * dz + dw * dw) / (2.0 * lambda * lambda)) * T * lambda;
        }
        for (int c = 0; c < 4; ++c) L[c] += (double)q[G_AR + c] * inner;
    }
    memcpy(out, L, sizeof(L));
}

/* The formulas in double on the fp32 direction as given (|n| = 1 only to fp32 rounding) ... */
ORC_API void orc_radiance_f64(const float *o, const float *n, const float *g, uint64_t count, int variant, double *out)
{
    const double nd[4] = {n[0], n[1], n[2], n[3]};
    radiance_f64_d(o, nd, g, count, variant, out);
}

/* ... and on the direction re-normalised in double: the closed form assumes |n| = 1 exactly (its
 * oc_sqnorm - mu_bar^2 is the squared ray-centre distance only then), so for small sigma at large depth the
 * fp32 normalisation error of n alone moves the result by O(|oc|^2 eps / sigma^2).  This entry is the
 * arbiter for such scenes (DESIGN.md "numerical conditioning"). */
ORC_API void orc_radiance_f64_unit(const float *o, const float *n, const float *g, uint64_t count, int variant, double *out)
{
    double nd[4] = {n[0], n[1], n[2], n[3]};
    const double len = sqrt(nd[0] * nd[0] + nd[1] * nd[1] + nd[2] * nd[2] + nd[3] * nd[3]);
    for (int i = 0; i < 4; ++i) nd[i] /= len;
    radiance_f64_d(o, nd, g, count, variant, out);
}

/* Batched: rays x one list, spread over host threads (pthreads; no OpenMP runtime in the image).
 * dirs: n_rays x 4 unit vectors (the reference's own fp32 directions). */
#include <pthread.h>
#include <unistd.h>

typedef struct
{
    const float *o, *dirs, *g;
    uint64_t n_rays, count;
    int variant, is64, tid, nthreads;
    void *out;
} ray_job_t;

static void *ray_worker(void *p)
{
    const ray_job_t *j = (const ray_job_t *)p;
    for (uint64_t r = (uint64_t)j->tid; r < j->n_rays; r += (uint64_t)j->nthreads)
    {
        if (j->is64 == 2) orc_radiance_f64_unit(j->o, j->dirs + 4 * r, j->g, j->count, j->variant, (double *)j->out + 4 * r);
        else if (j->is64) orc_radiance_f64(j->o, j->dirs + 4 * r, j->g, j->count, j->variant, (double *)j->out + 4 * r);
        else orc_radiance_f32(j->o, j->dirs + 4 * r, j->g, j->count, j->variant, (float *)j->out + 4 * r);
    }
    return NULL;
}

static void run_rays(const float *o, const float *dirs, uint64_t n_rays, const float *g, uint64_t count, int variant, int is64, void *out)
{
    long nt = sysconf(_SC_NPROCESSORS_ONLN);
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    if ((uint64_t)nt > n_rays) nt = n_rays ? (long)n_rays : 1;
    pthread_t th[64];
    ray_job_t jobs[64];
    for (long t = 0; t < nt; ++t)
    {
        jobs[t] = (ray_job_t){o, dirs, g, n_rays, count, variant, is64, (int)t, (int)nt, out};
        pthread_create(&th[t], NULL, ray_worker, &jobs[t]);
    }
    for (long t = 0; t < nt; ++t) pthread_join(th[t], NULL);
}

ORC_API void orc_radiance_rays_f32(const float *o, const float *dirs, uint64_t n_rays, const float *g, uint64_t count, int variant, float *out)
{
    run_rays(o, dirs, n_rays, g, count, variant, 0, out);
}

ORC_API void orc_radiance_rays_f64(const float *o, const float *dirs, uint64_t n_rays, const float *g, uint64_t count, int variant, double *out)
{
    run_rays(o, dirs, n_rays, g, count, variant, 1, out);
}

ORC_API void orc_radiance_rays_f64_unit(const float *o, const float *dirs, uint64_t n_rays, const float *g, uint64_t count, int variant, double *out)
{
    run_rays(o, dirs, n_rays, g, count, variant, 2, out);
}

/* ---------------------------------------------------------------- camera -------------------- */

static void normalize3(float *v)
{
    const float inv = 1.f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
}
static void cross3(const float *a, const float *b, float *r)
{
    r[0] = a[1] * b[2] - b[1] * a[2];
    r[1] = a[2] * b[0] - b[2] * a[0];
    r[2] = a[0] * b[1] - b[0] * a[1];
}
static float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* camera_t::turn + camera_t::update (src/vrt/camera.cpp:7-23, :52):
 *   front = normalize(cos(yaw)cos(pitch), sin(pitch), sin(yaw)cos(pitch)), pitch clamped to +-89 deg
 *   right = normalize(front x world_up), up = normalize(right x front)
 *   view  = translate(lookAt_RH(pos, pos + front, up), focal * front)
 * view16 is column-major (GLM layout). */
ORC_API void orc_view_matrix(const float *pos3, float yaw_deg, float pitch_deg, float focal, float *view16)
{
    const float d2r = 0.01745329251994329576923690768489f;
    if (pitch_deg > 89.f) pitch_deg = 89.f;
    if (pitch_deg < -89.f) pitch_deg = -89.f;
    float front[3] = {cosf(yaw_deg * d2r) * cosf(pitch_deg * d2r), sinf(pitch_deg * d2r), sinf(yaw_deg * d2r) * cosf(pitch_deg * d2r)};
    normalize3(front);
    const float world_up[3] = {0.f, 1.f, 0.f};
    float right[3], up[3];
    cross3(front, world_up, right); normalize3(right);
    cross3(right, front, up); normalize3(up);
    /* lookAt_RH(eye = pos, center = pos + front, up) */
    float f[3] = {(pos3[0] + front[0]) - pos3[0], (pos3[1] + front[1]) - pos3[1], (pos3[2] + front[2]) - pos3[2]};
    normalize3(f);
    float s[3], u[3];
    cross3(f, up, s); normalize3(s);
    cross3(s, f, u);
    float m[16] = {s[0], u[0], -f[0], 0.f, s[1], u[1], -f[1], 0.f, s[2], u[2], -f[2], 0.f, -dot3(s, pos3), -dot3(u, pos3), dot3(f, pos3), 1.f};
    /* translate(m, v): column3 = m0*v.x + m1*v.y + m2*v.z + m3 */
    const float v[3] = {focal * front[0], focal * front[1], focal * front[2]};
    for (int i = 0; i < 4; ++i) m[12 + i] = m[0 + i] * v[0] + m[4 + i] * v[1] + m[8 + i] * v[2] + m[12 + i];
    memcpy(view16, m, sizeof(m));
}

/* Orbit of the app camera, src/volumetric-ray-tracer/main.cpp:248-255: position (0,0,offset) rotated by
 * initial_rot about +y, yaw = -90 - initial_rot.  Outputs view (column-major) and origin (w = 0). */
ORC_API void orc_app_camera(float camera_offset, float focal, float initial_rot_deg, float *view16, float *origin4)
{
    const float a = initial_rot_deg * 0.01745329251994329576923690768489f;
    const float c = cosf(a), s = sinf(a);
    /* rotate(I, a, +y) * (0, 0, off, 1) = (sin(a)*off, 0, cos(a)*off) */
    float pos[3] = {s * camera_offset, 0.f, c * camera_offset};
    orc_view_matrix(pos, -90.f - initial_rot_deg, 0.f, focal, view16);
    origin4[0] = pos[0]; origin4[1] = pos[1]; origin4[2] = pos[2]; origin4[3] = 0.f;
}

/* General 4x4 inverse (adjugate / determinant); column-major in and out. */
ORC_API void orc_inverse4(const float *m, float *r)
{
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

/* Ray directions for pixel ids (row-major, id = row*w + col):
 *   plane = inverse(view) * (-1 + col/(w/2), -1 + row/(h/2), 0, 1)   src/vrt/camera.cpp:60-70
 *   dir   = normalize(plane.xyz0 - origin)  (all four lanes)         src/vrt/rt.h:231-236, types.h:75-82 */
ORC_API void orc_pixel_dirs(const float *view16, const float *origin4, uint64_t w, uint64_t h, const uint64_t *pix, uint64_t n_pix, float *dirs_out)
{
    float inv[16];
    orc_inverse4(view16, inv);
    for (uint64_t k = 0; k < n_pix; ++k)
    {
        const uint64_t row = pix[k] / w, col = pix[k] % w;
        const float x = -1.f + col / (w / 2.f), y = -1.f + row / (h / 2.f);
        float d[4];
        for (int i = 0; i < 3; ++i) d[i] = (inv[0 + i] * x + inv[4 + i] * y) + (inv[8 + i] * 0.f + inv[12 + i] * 1.f) - origin4[i];
        d[3] = 0.f - origin4[3];
        const float norm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3]);
        for (int i = 0; i < 4; ++i) dirs_out[4 * k + i] = d[i] / norm;
    }
}

/* ---------------------------------------------------------------- tiling -------------------- */

/* vrt::tile_gaussians, src/vrt/rt.cpp:29-69.  Writes per-tile counts (row-major, y outer) and the
 * concatenated index lists; returns the total number of entries (idx_out truncated at idx_cap), or
 * UINT64_MAX when counts_cap is too small.  Tile centres come from the same float accumulation
 * loops as the reference (rt.cpp:47-49); *tiles_w/_h follow tiles_t (types.h:280). */
ORC_API uint64_t orc_tile_membership(float tw, float th, const float *g, uint64_t n, const float *view16,
                                     uint64_t *tiles_w, uint64_t *tiles_h, uint64_t *n_lists, uint32_t *counts_out, uint64_t counts_cap,
                                     uint32_t *idx_out, uint64_t idx_cap)
{
    float *pmx = (float *)malloc(sizeof(float) * (n + 1)), *pmy = (float *)malloc(sizeof(float) * (n + 1));
    float *psg = (float *)malloc(sizeof(float) * (n + 1));
    uint32_t *ids = (uint32_t *)malloc(sizeof(uint32_t) * (n + 1));
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; ++i)
    {
        const float *q = g + i * G_STRIDE;
        const float x = q[G_MX], y = q[G_MY], z = q[G_MZ];
        /* proj = view * (mu.xyz, 1), GLM order (c0*x + c1*y) + (c2*z + c3*w) */
        const float px = (view16[0] * x + view16[4] * y) + (view16[8] * z + view16[12] * 1.f);
        const float py = (view16[1] * x + view16[5] * y) + (view16[9] * z + view16[13] * 1.f);
        const float pz = (view16[2] * x + view16[6] * y) + (view16[10] * z + view16[14] * 1.f);
        if (pz < 1.f) continue;
        const float sg = q[G_SIGMA] / pz;
        if (sg < 1e-5f) continue;
        pmx[m] = px / pz; pmy[m] = py / pz; psg[m] = sg; ids[m] = (uint32_t)i; ++m;
    }
    uint64_t total = 0, t = 0, nx = 0, ny = 0;
    for (float y = -1.f + th / 2; y < 1.f; y += th)
    {
        ++ny; nx = 0;
        for (float x = -1.f + tw / 2; x < 1.f; x += tw)
        {
            ++nx;
            uint32_t c = 0;
            for (uint64_t i = 0; i < m; ++i)
            {
                const float dx = fabsf(x - pmx[i]), dy = fabsf(y - pmy[i]);
                if (dx <= fabsf(x) + tw / 2 + 3.3f * psg[i] && dy <= fabsf(y) + th / 2 + 3.3f * psg[i])
                {
                    if (total < idx_cap) idx_out[total] = ids[i];
                    ++total; ++c;
                }
            }
            if (t < counts_cap) counts_out[t] = c;
            ++t;
        }
    }
    free(pmx); free(pmy); free(psg); free(ids);
    *tiles_w = (uint64_t)ceilf(2.f / tw);
    *tiles_h = (uint64_t)ceilf(2.f / th);
    (void)nx; (void)ny;
    *n_lists = t;
    return t > counts_cap ? UINT64_MAX : total;
}

/* ---------------------------------------------------------------- framebuffer --------------- */

/* Pixel packing.  quantise 0: truncate, alpha 0xFF (scalar paths, rt.h:238-243, 278-283);
 * quantise 1: round to nearest even (simd::cvts, rt.h:329-333); alpha_quirk: tiled SIMD path
 * writes A = min(1, color.w)*255 (rt.h:373-377). */
ORC_API uint32_t orc_pack_pixel(const float *rgba, int round_nearest, int alpha_quirk)
{
    uint32_t ch[4];
    for (int c = 0; c < 4; ++c)
    {
        const float v = fminf(rgba[c], 1.0f) * 255.f;
        ch[c] = round_nearest ? (uint32_t)(int32_t)nearbyintf(v) : (uint32_t)v;
    }
    const uint32_t A = alpha_quirk ? (ch[3] << 24) : 0xFF000000u;
    return A | ch[0] << 16 | ch[1] << 8 | ch[2];
}

/* ---------------------------------------------------------------- OBJ ingestion ------------- */

/* read_from_obj, src/vrt/gaussians-from-file.cpp:7-44: every `v x y z` becomes a Gaussian with
 * sigma by vertex count (<300: 0.3, <1000: 0.15, else 0.05), albedo = 0.5*normalize(p) + 0.5
 * (w = 1), magnitude 1.  Returns the vertex count (copies up to cap), UINT64_MAX on I/O error. */
ORC_API uint64_t orc_read_obj(const char *path, float *aos_out, uint64_t cap)
{
    FILE *f = fopen(path, "r");
    if (!f) return UINT64_MAX;
    char line[1024];
    uint64_t n = 0;
    while (fgets(line, sizeof(line), f))
    {
        float x, y, z;
        if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t') && sscanf(line + 2, "%f %f %f", &x, &y, &z) == 3)
        {
            if (n < cap)
            {
                float *q = aos_out + n * G_STRIDE;
                const float norm = sqrtf(x * x + y * y + z * z + 0.f);
                q[G_AR] = (x / norm) * 0.5f + 0.5f; q[G_AG] = (y / norm) * 0.5f + 0.5f; q[G_AB] = (z / norm) * 0.5f + 0.5f;
                q[G_AW] = (0.f / norm) * 0.5f + 1.0f;
                q[G_MX] = x; q[G_MY] = y; q[G_MZ] = z; q[G_MW] = 0.f;
                q[G_MAG] = 1.0f;
            }
            ++n;
        }
    }
    fclose(f);
    const float sig = n < 300 ? 0.3f : (n < 1000 ? 0.15f : 0.05f);
    for (uint64_t i = 0; i < n && i < cap; ++i) aos_out[i * G_STRIDE + G_SIGMA] = sig;
    return n;
}
