// ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing in the product path (strelka_b200/) may include,
// link or call anything under oracle/.  See oracle/README.md.
//
// Small float vector helpers for the CPU restatement of Strelka's path tracer.
//
// Arithmetic contract (shared *by specification*, not by source, with the CUDA kernels): the
// oracle is compiled with -ffp-contract=off, so every expression below is evaluated exactly
// as written in IEEE-754 binary32; fused multiply-adds appear only where written as fmaf().
// The CUDA kernels are compiled with -fmad=false and spell out the same expression trees, so the
// two sides differ only through libm (sin/cos/acos/pow) -- see DESIGN.md "Arithmetic contract".
// dot()/cross() follow sutil/vec_math.h literally (unfused), which is also what a host build of the
// reference headers evaluates, so the header-level golden vectors are matched bit for bit.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace orc
{

struct f2
{
    float x, y;
};
struct f3
{
    float x, y, z;
};
struct f4
{
    float x, y, z, w;
};

inline f3 mk3(float x, float y, float z)
{
    return f3{ x, y, z };
}
inline f3 mk3(float s)
{
    return f3{ s, s, s };
}
inline f3 mk3(const f4& v)
{
    return f3{ v.x, v.y, v.z };
}
inline f4 mk4(const f3& v, float w)
{
    return f4{ v.x, v.y, v.z, w };
}

inline f3 operator+(const f3& a, const f3& b)
{
    return f3{ a.x + b.x, a.y + b.y, a.z + b.z };
}
inline f3 operator-(const f3& a, const f3& b)
{
    return f3{ a.x - b.x, a.y - b.y, a.z - b.z };
}
inline f3 operator-(const f3& a)
{
    return f3{ -a.x, -a.y, -a.z };
}
inline f3 operator*(const f3& a, const f3& b)
{
    return f3{ a.x * b.x, a.y * b.y, a.z * b.z };
}
inline f3 operator*(const f3& a, float s)
{
    return f3{ a.x * s, a.y * s, a.z * s };
}
inline f3 operator*(float s, const f3& a)
{
    return f3{ s * a.x, s * a.y, s * a.z };
}
inline f3 operator/(const f3& a, float s)
{
    // sutil/vec_math.h:478-482: multiply by the reciprocal
    const float inv = 1.0f / s;
    return f3{ a.x * inv, a.y * inv, a.z * inv };
}
inline f3 operator/(const f3& a, const f3& b)
{
    return f3{ a.x / b.x, a.y / b.y, a.z / b.z };
}
inline f3& operator+=(f3& a, const f3& b)
{
    a = a + b;
    return a;
}
inline f3& operator*=(f3& a, const f3& b)
{
    a = a * b;
    return a;
}
inline f3& operator*=(f3& a, float s)
{
    a = a * s;
    return a;
}

inline f4 operator+(const f4& a, const f4& b)
{
    return f4{ a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w };
}
inline f4 operator-(const f4& a, const f4& b)
{
    return f4{ a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w };
}
inline f4 operator*(const f4& a, float s)
{
    return f4{ a.x * s, a.y * s, a.z * s, a.w * s };
}
inline f4 operator*(float s, const f4& a)
{
    return f4{ s * a.x, s * a.y, s * a.z, s * a.w };
}
inline f4 operator/(const f4& a, float s)
{
    const float inv = 1.0f / s;
    return f4{ a.x * inv, a.y * inv, a.z * inv, a.w * inv };
}

// sutil/vec_math.h:530-539, evaluated left to right without fusing (what a host build of the
// reference headers computes; oracle/ref_crosscheck.cpp pins this bit for bit)
inline float dot(const f3& a, const f3& b)
{
    return a.x * b.x + a.y * b.y + a.z * b.z;
}
inline f3 cross(const f3& a, const f3& b)
{
    return f3{ a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
}
// This repository's own geometric kernels (triangle test, instance flattening) have no reference
// counterpart (OptiX does them); they are DEFINED with these fused forms on both sides.
inline float dot_fma(const f3& a, const f3& b)
{
    return std::fmaf(a.z, b.z, std::fmaf(a.y, b.y, a.x * b.x));
}
inline f3 cross_fma(const f3& a, const f3& b)
{
    return f3{ std::fmaf(a.y, b.z, -(a.z * b.y)), std::fmaf(a.z, b.x, -(a.x * b.z)), std::fmaf(a.x, b.y, -(a.y * b.x)) };
}
inline float length(const f3& v)
{
    return std::sqrt(dot(v, v));
}
// sutil/vec_math.h:549-553: v * (1/sqrt(dot))
inline f3 normalize(const f3& v)
{
    const float inv = 1.0f / std::sqrt(dot(v, v));
    return v * inv;
}
inline float clampf(float v, float lo, float hi)
{
    return std::fmax(lo, std::fmin(v, hi));
}
inline float saturate(float v)
{
    return clampf(v, 0.0f, 1.0f);
}
// sutil/vec_math.h:504-507: a + t*(b-a)
inline f3 lerp(const f3& a, const f3& b, float t)
{
    return a + t * (b - a);
}
inline bool all_nonzero(const f3& v)
{
    return v.x != 0.0f && v.y != 0.0f && v.z != 0.0f;
}
inline bool isnan3(const f3& v)
{
    return std::isnan(v.x) || std::isnan(v.y) || std::isnan(v.z);
}
inline float maxcomp(const f3& v)
{
    return std::fmax(v.x, std::fmax(v.y, v.z));
}
inline uint32_t f2u(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float u2f(uint32_t u)
{
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
inline int32_t f2i(float f)
{
    int32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float i2f(int32_t u)
{
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// sin(2*pi*u), cos(2*pi*u), u in [0,1): the fixed-polynomial routine shared BY SPECIFICATION with the
// CUDA kernels (strelka_b200/csrc/hd.cuh) so that BSDF sampling yields identical bits on both sides.
inline void sincos2pi(float u, float& s, float& c)
{
    const float q = std::floor(std::fmaf(u, 4.0f, 0.5f));
    const float r = std::fmaf(q, -0.25f, u);
    const float x = r * 6.283185307179586f;
    const float x2 = x * x;
    float sp = std::fmaf(x2, 2.7557319e-6f, -1.9841270e-4f);
    sp = std::fmaf(sp, x2, 8.3333333e-3f);
    sp = std::fmaf(sp, x2, -1.6666667e-1f);
    sp = std::fmaf(sp * x2, x, x);
    float cp = std::fmaf(x2, -2.7557319e-7f, 2.4801587e-5f);
    cp = std::fmaf(cp, x2, -1.3888889e-3f);
    cp = std::fmaf(cp, x2, 4.1666667e-2f);
    cp = std::fmaf(cp, x2, -0.5f);
    cp = std::fmaf(cp, x2, 1.0f);
    const int k = int(q) & 3;
    s = (k == 0) ? sp : (k == 1) ? cp : (k == 2) ? -sp : -cp;
    c = (k == 0) ? cp : (k == 1) ? -sp : (k == 2) ? -cp : sp;
}

// Affine 3x4 transform (rows), applied with the fixed fma chain
//   r = fma(m0,x, fma(m1,y, fma(m2,z, m3)))
struct Affine
{
    float m[12]; // row-major 3x4
};
inline f3 xform_point(const Affine& a, const f3& p)
{
    return f3{ std::fmaf(a.m[0], p.x, std::fmaf(a.m[1], p.y, std::fmaf(a.m[2], p.z, a.m[3]))),
               std::fmaf(a.m[4], p.x, std::fmaf(a.m[5], p.y, std::fmaf(a.m[6], p.z, a.m[7]))),
               std::fmaf(a.m[8], p.x, std::fmaf(a.m[9], p.y, std::fmaf(a.m[10], p.z, a.m[11]))) };
}
inline f3 xform_vector(const Affine& a, const f3& v)
{
    return f3{ std::fmaf(a.m[0], v.x, std::fmaf(a.m[1], v.y, a.m[2] * v.z)),
               std::fmaf(a.m[4], v.x, std::fmaf(a.m[5], v.y, a.m[6] * v.z)),
               std::fmaf(a.m[8], v.x, std::fmaf(a.m[9], v.y, a.m[10] * v.z)) };
}
// normal transform = transpose(inverse) applied to n: rows of inv become columns
// (optixTransformNormalFromObjectToWorldSpace semantic): n' = inv^T * n
inline f3 xform_normal(const Affine& inv, const f3& n)
{
    return f3{ std::fmaf(inv.m[0], n.x, std::fmaf(inv.m[4], n.y, inv.m[8] * n.z)),
               std::fmaf(inv.m[1], n.x, std::fmaf(inv.m[5], n.y, inv.m[9] * n.z)),
               std::fmaf(inv.m[2], n.x, std::fmaf(inv.m[6], n.y, inv.m[10] * n.z)) };
}

} // namespace orc
