// Common host/device helpers for the B200 backend.
//
// All per-element device logic in this directory is written as SB_HD functions so that, besides
// being compiled by nvcc for sm_100a (the product), the very same bodies can be instantiated by
// g++ inside tests/emul (a serial host harness used ONLY by the CPU test-suite to catch logic bugs
// before a GPU run).  The product library contains no CPU execution path.
//
// Arithmetic contract: device code is compiled with -fmad=false, so expressions are
// evaluated as written; fused multiply-adds appear only as explicit fmaf().  dot()/cross() are the
// unfused sutil forms the reference uses (sutil/vec_math.h:530-539); dot_fma()/cross_fma() are this
// backend's own forms for code with no reference counterpart (triangle test, flattening, traversal).
#pragma once

#include <cstdint>
#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define SB_HD __host__ __device__ __forceinline__
#define SB_D __device__ __forceinline__
#else
#include <vector_types.h> // float3/float4/uint2/uint4 PODs from the CUDA toolkit headers
#define SB_HD inline
#define SB_D inline
#endif

namespace sb
{

constexpr float kPi = 3.14159265358979323846f;
constexpr float kOneMinusEps = 0x1.fffffep-1f;

// ---- bit helpers ---------------------------------------------------------------------------------
SB_HD uint32_t f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
SB_HD float u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
SB_HD uint32_t brev32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __brev(v);
#else
    v = ((v & 0xaaaaaaaau) >> 1) | ((v & 0x55555555u) << 1);
    v = ((v & 0xccccccccu) >> 2) | ((v & 0x33333333u) << 2);
    v = ((v & 0xf0f0f0f0u) >> 4) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v & 0xff00ff00u) >> 8) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
#endif
}
SB_HD uint32_t popc32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return uint32_t(__builtin_popcount(v));
#endif
}
// index of the highest set bit (v != 0)
SB_HD uint32_t bfind32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return 31u - uint32_t(__clz(int(v)));
#else
    return 31u - uint32_t(__builtin_clz(v));
#endif
}
SB_HD float fminf_(float a, float b)
{
    return fminf(a, b);
}
SB_HD float fmaxf_(float a, float b)
{
    return fmaxf(a, b);
}

// ---- float3 math ---------------------------------------------------------------------------------
SB_HD float3 mk3(float x, float y, float z)
{
    float3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
SB_HD float3 mk3(float s)
{
    return mk3(s, s, s);
}
SB_HD float3 mk3(const float4& v)
{
    return mk3(v.x, v.y, v.z);
}
SB_HD float4 mk4(float x, float y, float z, float w)
{
    float4 r;
    r.x = x;
    r.y = y;
    r.z = z;
    r.w = w;
    return r;
}
SB_HD float4 mk4(const float3& v, float w)
{
    return mk4(v.x, v.y, v.z, w);
}
SB_HD float3 operator+(const float3& a, const float3& b)
{
    return mk3(a.x + b.x, a.y + b.y, a.z + b.z);
}
SB_HD float3 operator-(const float3& a, const float3& b)
{
    return mk3(a.x - b.x, a.y - b.y, a.z - b.z);
}
SB_HD float3 operator-(const float3& a)
{
    return mk3(-a.x, -a.y, -a.z);
}
SB_HD float3 operator*(const float3& a, const float3& b)
{
    return mk3(a.x * b.x, a.y * b.y, a.z * b.z);
}
SB_HD float3 operator*(const float3& a, float s)
{
    return mk3(a.x * s, a.y * s, a.z * s);
}
SB_HD float3 operator*(float s, const float3& a)
{
    return mk3(s * a.x, s * a.y, s * a.z);
}
// sutil/vec_math.h:486-490: divide = multiply by the reciprocal
SB_HD float3 operator/(const float3& a, float s)
{
    const float inv = 1.0f / s;
    return mk3(a.x * inv, a.y * inv, a.z * inv);
}
SB_HD float3 operator/(const float3& a, const float3& b)
{
    return mk3(a.x / b.x, a.y / b.y, a.z / b.z);
}
SB_HD float3& operator+=(float3& a, const float3& b)
{
    a = a + b;
    return a;
}
SB_HD float3& operator*=(float3& a, const float3& b)
{
    a = a * b;
    return a;
}
SB_HD float3& operator*=(float3& a, float s)
{
    a = a * s;
    return a;
}
SB_HD float4 operator+(const float4& a, const float4& b)
{
    return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
SB_HD float4 operator-(const float4& a, const float4& b)
{
    return mk4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}
SB_HD float4 operator*(const float4& a, float s)
{
    return mk4(a.x * s, a.y * s, a.z * s, a.w * s);
}
SB_HD float4 operator*(float s, const float4& a)
{
    return mk4(s * a.x, s * a.y, s * a.z, s * a.w);
}
SB_HD float4 operator/(const float4& a, float s)
{
    const float inv = 1.0f / s;
    return mk4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
}
SB_HD float dot(const float3& a, const float3& b)
{
    return a.x * b.x + a.y * b.y + a.z * b.z;
}
SB_HD float3 cross(const float3& a, const float3& b)
{
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
SB_HD float dot_fma(const float3& a, const float3& b)
{
    return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x));
}
SB_HD float3 cross_fma(const float3& a, const float3& b)
{
    return mk3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
SB_HD float length(const float3& v)
{
    return sqrtf(dot(v, v));
}
SB_HD float3 normalize(const float3& v)
{
    const float inv = 1.0f / sqrtf(dot(v, v));
    return v * inv;
}
SB_HD float clampf(float v, float lo, float hi)
{
    return fmaxf(lo, fminf(v, hi));
}
SB_HD float saturate(float v)
{
    return clampf(v, 0.0f, 1.0f);
}
SB_HD float3 lerp(const float3& a, const float3& b, float t)
{
    return a + t * (b - a);
}
SB_HD bool all_nonzero(const float3& v)
{
    return v.x != 0.0f && v.y != 0.0f && v.z != 0.0f;
}
SB_HD bool isnanf_(float v)
{
    return v != v;
}
SB_HD bool isnan3(const float3& v)
{
    return isnanf_(v.x) || isnanf_(v.y) || isnanf_(v.z);
}
SB_HD float maxcomp(const float3& v)
{
    return fmaxf(v.x, fmaxf(v.y, v.z));
}

// sin(2*pi*u), cos(2*pi*u) for u in [0,1): quadrant reduction in turns + fixed fmaf polynomials.
// libm's sinf/cosf differ by an ulp or two between CUDA and glibc, and on hair-thin geometry that is
// enough to flip hits; the BSDF samplers (whose definition is ours anyway) therefore use this routine,
// which evaluates to the same bits on the device and in the CPU oracle (max error ~1 ulp).
SB_HD void sincos2pi(float u, float& s, float& c)
{
    const float q = floorf(fmaf(u, 4.0f, 0.5f)); // nearest quarter turn: 0..4
    const float r = fmaf(q, -0.25f, u); // |r| <= 1/8 turn, exact
    const float x = r * 6.283185307179586f; // |x| <= pi/4
    const float x2 = x * x;
    float sp = fmaf(x2, 2.7557319e-6f, -1.9841270e-4f); // x^9/9!, x^7/7!
    sp = fmaf(sp, x2, 8.3333333e-3f);
    sp = fmaf(sp, x2, -1.6666667e-1f);
    sp = fmaf(sp * x2, x, x);
    float cp = fmaf(x2, -2.7557319e-7f, 2.4801587e-5f); // x^10/10!, x^8/8!
    cp = fmaf(cp, x2, -1.3888889e-3f);
    cp = fmaf(cp, x2, 4.1666667e-2f);
    cp = fmaf(cp, x2, -0.5f);
    cp = fmaf(cp, x2, 1.0f);
    const int k = int(q) & 3;
    s = (k == 0) ? sp : (k == 1) ? cp : (k == 2) ? -sp : -cp;
    c = (k == 0) ? cp : (k == 1) ? -sp : (k == 2) ? -cp : sp;
}

// ---- affine transforms (row-major 3x4) ---------------------------------------------------------------
struct Affine
{
    float m[12];
};
SB_HD float3 xform_point(const Affine& a, const float3& p)
{
    return mk3(fmaf(a.m[0], p.x, fmaf(a.m[1], p.y, fmaf(a.m[2], p.z, a.m[3]))),
               fmaf(a.m[4], p.x, fmaf(a.m[5], p.y, fmaf(a.m[6], p.z, a.m[7]))),
               fmaf(a.m[8], p.x, fmaf(a.m[9], p.y, fmaf(a.m[10], p.z, a.m[11]))));
}
// n' = inv^T * n  (optixTransformNormalFromObjectToWorldSpace)
SB_HD float3 xform_normal(const Affine& inv, const float3& n)
{
    return mk3(fmaf(inv.m[0], n.x, fmaf(inv.m[4], n.y, inv.m[8] * n.z)), fmaf(inv.m[1], n.x, fmaf(inv.m[5], n.y, inv.m[9] * n.z)),
               fmaf(inv.m[2], n.x, fmaf(inv.m[6], n.y, inv.m[10] * n.z)));
}

} // namespace sb
