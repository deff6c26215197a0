// Minimal GLM stand-in, written from scratch for this repo (MIT-style, no GLM code).
//
// TEST INFRASTRUCTURE ONLY.  GLM is neither installed in the build image nor
// vendored by the reference, yet the reference's vrt sources include it
// (src/vrt/camera.cpp:15-21,52,64; src/vrt/rt.cpp:37-57; src/vrt/types.h:84-91).
// This header supplies exactly the subset those files use so that the
// UNMODIFIED reference sources can be compiled in place into oracle/_ref/.
// Semantics follow the published GLM conventions: column-major mat4,
// m[c] is column c, right-handed lookAt, translate() post-multiplies.
#pragma once
#include <cmath>

namespace glm
{
    struct vec2
    {
        float x, y;
        vec2() : x(0.f), y(0.f) {}
        vec2(float x_, float y_) : x(x_), y(y_) {}
    };
    inline vec2 operator-(const vec2 &a, const vec2 &b) { return vec2(a.x - b.x, a.y - b.y); }
    inline vec2 operator+(const vec2 &a, const vec2 &b) { return vec2(a.x + b.x, a.y + b.y); }
    inline vec2 abs(const vec2 &a) { return vec2(std::fabs(a.x), std::fabs(a.y)); }

    struct vec4;
    struct vec3
    {
        float x, y, z;
        vec3() : x(0.f), y(0.f), z(0.f) {}
        explicit vec3(float s) : x(s), y(s), z(s) {}
        vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
        vec3(const vec4 &v); // drops w, as GLM's truncating constructor does
        float &operator[](int i) { return (&x)[i]; }
        const float &operator[](int i) const { return (&x)[i]; }
    };
    inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
    inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
    inline vec3 operator-(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }
    inline vec3 operator*(const vec3 &a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
    inline vec3 operator*(float s, const vec3 &a) { return vec3(a.x * s, a.y * s, a.z * s); }
    inline float dot(const vec3 &a, const vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
    inline vec3 cross(const vec3 &a, const vec3 &b)
    {
        return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
    }
    inline vec3 normalize(const vec3 &a)
    {
        const float inv = 1.f / std::sqrt(dot(a, a));
        return a * inv;
    }

    struct vec4
    {
        float x, y, z, w;
        vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
        explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
        vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
        vec4(const vec3 &v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
        float &operator[](int i) { return (&x)[i]; }
        const float &operator[](int i) const { return (&x)[i]; }
    };
    inline vec3::vec3(const vec4 &v) : x(v.x), y(v.y), z(v.z) {}
    inline vec4 operator+(const vec4 &a, const vec4 &b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
    inline vec4 operator-(const vec4 &a, const vec4 &b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
    inline vec4 operator*(const vec4 &a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
    inline vec4 operator*(float s, const vec4 &a) { return a * s; }

    inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

    struct mat4
    {
        vec4 c[4]; // columns
        mat4() : mat4(1.f) {}
        explicit mat4(float d)
        {
            c[0] = vec4(d, 0.f, 0.f, 0.f);
            c[1] = vec4(0.f, d, 0.f, 0.f);
            c[2] = vec4(0.f, 0.f, d, 0.f);
            c[3] = vec4(0.f, 0.f, 0.f, d);
        }
        vec4 &operator[](int i) { return c[i]; }
        const vec4 &operator[](int i) const { return c[i]; }
    };

    inline vec4 operator*(const mat4 &m, const vec4 &v)
    {
        // (c0*x + c1*y) + (c2*z + c3*w)
        return (m[0] * v.x + m[1] * v.y) + (m[2] * v.z + m[3] * v.w);
    }
    inline mat4 operator*(const mat4 &a, const mat4 &b)
    {
        mat4 r(0.f);
        for (int j = 0; j < 4; ++j) r[j] = a * b[j];
        return r;
    }

    inline mat4 translate(const mat4 &m, const vec3 &v)
    {
        mat4 r = m;
        r[3] = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3];
        return r;
    }

    /// Rotation by `angle` (radians) about `axis`, post-multiplied onto `m`.
    inline mat4 rotate(const mat4 &m, float angle, const vec3 &axis_in)
    {
        const float c = std::cos(angle), s = std::sin(angle);
        const vec3 a = normalize(axis_in);
        const vec3 t = a * (1.f - c);
        float R[3][3];
        R[0][0] = c + t.x * a.x;       R[0][1] = t.x * a.y + s * a.z; R[0][2] = t.x * a.z - s * a.y;
        R[1][0] = t.y * a.x - s * a.z; R[1][1] = c + t.y * a.y;       R[1][2] = t.y * a.z + s * a.x;
        R[2][0] = t.z * a.x + s * a.y; R[2][1] = t.z * a.y - s * a.x; R[2][2] = c + t.z * a.z;
        mat4 r(0.f);
        for (int j = 0; j < 3; ++j) r[j] = m[0] * R[j][0] + m[1] * R[j][1] + m[2] * R[j][2];
        r[3] = m[3];
        return r;
    }

    /// Right-handed look-at (GLM default clip-space configuration).
    inline mat4 lookAt(const vec3 &eye, const vec3 &center, const vec3 &up)
    {
        const vec3 f = normalize(center - eye);
        const vec3 s = normalize(cross(f, up));
        const vec3 u = cross(s, f);
        mat4 r(1.f);
        r[0][0] = s.x;  r[1][0] = s.y;  r[2][0] = s.z;
        r[0][1] = u.x;  r[1][1] = u.y;  r[2][1] = u.z;
        r[0][2] = -f.x; r[1][2] = -f.y; r[2][2] = -f.z;
        r[3][0] = -dot(s, eye);
        r[3][1] = -dot(u, eye);
        r[3][2] = dot(f, eye);
        return r;
    }

    /// General 4x4 inverse by cofactor expansion (adjugate / determinant).
    inline mat4 inverse(const mat4 &m)
    {
        const float a00 = m[0][0], a01 = m[0][1], a02 = m[0][2], a03 = m[0][3];
        const float a10 = m[1][0], a11 = m[1][1], a12 = m[1][2], a13 = m[1][3];
        const float a20 = m[2][0], a21 = m[2][1], a22 = m[2][2], a23 = m[2][3];
        const float a30 = m[3][0], a31 = m[3][1], a32 = m[3][2], a33 = m[3][3];

        const float b00 = a00 * a11 - a01 * a10, b01 = a00 * a12 - a02 * a10;
        const float b02 = a00 * a13 - a03 * a10, b03 = a01 * a12 - a02 * a11;
        const float b04 = a01 * a13 - a03 * a11, b05 = a02 * a13 - a03 * a12;
        const float b06 = a20 * a31 - a21 * a30, b07 = a20 * a32 - a22 * a30;
        const float b08 = a20 * a33 - a23 * a30, b09 = a21 * a32 - a22 * a31;
        const float b10 = a21 * a33 - a23 * a31, b11 = a22 * a33 - a23 * a32;

        const float det = b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06;
        const float id = 1.f / det;

        mat4 r(0.f);
        r[0][0] = (a11 * b11 - a12 * b10 + a13 * b09) * id;
        r[0][1] = (a02 * b10 - a01 * b11 - a03 * b09) * id;
        r[0][2] = (a31 * b05 - a32 * b04 + a33 * b03) * id;
        r[0][3] = (a22 * b04 - a21 * b05 - a23 * b03) * id;
        r[1][0] = (a12 * b08 - a10 * b11 - a13 * b07) * id;
        r[1][1] = (a00 * b11 - a02 * b08 + a03 * b07) * id;
        r[1][2] = (a32 * b02 - a30 * b05 - a33 * b01) * id;
        r[1][3] = (a20 * b05 - a22 * b02 + a23 * b01) * id;
        r[2][0] = (a10 * b10 - a11 * b08 + a13 * b06) * id;
        r[2][1] = (a01 * b08 - a00 * b10 - a03 * b06) * id;
        r[2][2] = (a30 * b04 - a31 * b02 + a33 * b00) * id;
        r[2][3] = (a21 * b02 - a20 * b04 - a23 * b00) * id;
        r[3][0] = (a11 * b07 - a10 * b09 - a12 * b06) * id;
        r[3][1] = (a00 * b09 - a01 * b07 + a02 * b06) * id;
        r[3][2] = (a31 * b01 - a30 * b03 - a32 * b00) * id;
        r[3][3] = (a20 * b03 - a21 * b01 + a22 * b00) * id;
        return r;
    }
}
