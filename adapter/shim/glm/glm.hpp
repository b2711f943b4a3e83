// MINIMAL STAND-IN for the handful of glm types the adapter touches, so that adapter/B200Render.cpp can
// be compiled and exercised in this repository where glm is not installed.  Inside a Strelka tree the
// real <glm/glm.hpp> is used instead (do not put adapter/shim on the include path there).
#pragma once
#include <cmath>
namespace glm
{
struct float2 { float x, y; };
struct float3 { float x, y, z; float3() : x(0), y(0), z(0) {} float3(float a, float b, float c) : x(a), y(b), z(c) {} };
struct float4
{
    float x, y, z, w;
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
struct quat { float w, x, y, z; };
// column-major 4x4: m[c][r]
struct mat4
{
    float4 c[4];
    mat4() : mat4(1.0f) {}
    explicit mat4(float d) { for (int i = 0; i < 4; ++i) c[i] = float4(i == 0 ? d : 0, i == 1 ? d : 0, i == 2 ? d : 0, i == 3 ? d : 0); }
    float4& operator[](int i) { return c[i]; }
    const float4& operator[](int i) const { return c[i]; }
};
using float4x4 = mat4;
inline const float* value_ptr(const mat4& m) { return &m.c[0].x; }
inline bool operator==(const mat4& a, const mat4& b)
{
    for (int i = 0; i < 16; ++i)
        if (value_ptr(a)[i] != value_ptr(b)[i])
            return false;
    return true;
}
} // namespace glm
