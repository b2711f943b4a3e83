// TEST INFRASTRUCTURE: a minimal stand-in for the subset of the glm 0.9.9 API (G-Truc Creation, MIT licence) that
// Strelka's own headers and scene/camera sources use -- glm itself is an external dependency of the reference
// (conanfile.py) that is not installed in this image.  With this directory on the include path the reference's REAL
// include/render, include/scene, include/settings headers and src/scene/{scene,camera}.cpp compile unmodified, so that
// adapter/B200Render.cpp can be checked against the real oka::Render / Buffer / Scene declarations and the Python scene
// mirror against the real oka::Scene flattening (tests/test_adapter_real_headers.py).  Written from glm's documented
// semantics (column-major matrices, column vectors, right-handed, quaternions stored x, y, z, w and constructed
// (w, x, y, z)); it is not a copy of glm and covers nothing beyond what those files reference.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace glm
{

struct vec2
{
    float x = 0.0f, y = 0.0f;
    vec2() = default;
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float x_, float y_) : x(x_), y(y_) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};

struct vec4;
struct vec3
{
    float x = 0.0f, y = 0.0f, z = 0.0f;
    vec3() = default;
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    vec3(double x_, double y_, double z_) : x(float(x_)), y(float(y_)), z(float(z_)) {}
    vec3(const vec4& v); // implicit, as in glm without GLM_FORCE_EXPLICIT_CTOR
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
    vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4
{
    float x = 0.0f, y = 0.0f, z = 0.0f, w = 0.0f;
    vec4() = default;
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    vec4(double x_, double y_, double z_, double w_) : x(float(x_)), y(float(y_)), z(float(z_)), w(float(w_)) {}
    vec4(const vec3& v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}

inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return a * s; }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(const vec3& a, const vec3& b) { return !(a == b); }
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator-(const vec4& a, const vec4& b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator*(float s, const vec4& a) { return a * s; }
inline vec4 operator/(const vec4& a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline bool operator==(const vec4& a, const vec4& b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }

inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(const vec3& a) { return std::sqrt(dot(a, a)); }
inline float length(const vec4& a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(const vec3& a) { return a * (1.0f / length(a)); }
inline vec4 normalize(const vec4& a) { return a * (1.0f / length(a)); }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
inline vec3 radians(const vec3& d) { return vec3(radians(d.x), radians(d.y), radians(d.z)); }
inline float degrees(float rad) { return rad * 57.295779513082320876798154814105f; }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(const vec3& a, const vec3& b, float t) { return a * (1.0f - t) + b * t; }
inline vec4 mix(const vec4& a, const vec4& b, float t) { return a * (1.0f - t) + b * t; }
inline float clamp(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)); }

// quaternion: members in glm's default storage order x, y, z, w; constructed (w, x, y, z)
struct quat
{
    float x = 0.0f, y = 0.0f, z = 0.0f, w = 1.0f;
    quat() = default;
    quat(float w_, float x_, float y_, float z_) : x(x_), y(y_), z(z_), w(w_) {}
    quat(float w_, const vec3& v) : x(v.x), y(v.y), z(v.z), w(w_) {}
    // from Euler angles (pitch = x, yaw = y, roll = z), radians
    explicit quat(const vec3& e)
    {
        const vec3 c(std::cos(e.x * 0.5f), std::cos(e.y * 0.5f), std::cos(e.z * 0.5f));
        const vec3 s(std::sin(e.x * 0.5f), std::sin(e.y * 0.5f), std::sin(e.z * 0.5f));
        w = c.x * c.y * c.z + s.x * s.y * s.z;
        x = s.x * c.y * c.z - c.x * s.y * s.z;
        y = c.x * s.y * c.z + s.x * c.y * s.z;
        z = c.x * c.y * s.z - s.x * s.y * c.z;
    }
};
inline quat operator*(const quat& p, const quat& q)
{
    return quat(p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z, p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y,
                p.w * q.y + p.y * q.w + p.z * q.x - p.x * q.z, p.w * q.z + p.z * q.w + p.x * q.y - p.y * q.x);
}
inline quat operator*(const quat& q, float s) { return quat(q.w * s, q.x * s, q.y * s, q.z * s); }
inline quat operator+(const quat& a, const quat& b) { return quat(a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z); }
inline quat operator-(const quat& q) { return quat(-q.w, -q.x, -q.y, -q.z); }
inline vec3 operator*(const quat& q, const vec3& v)
{
    const vec3 qv(q.x, q.y, q.z);
    const vec3 uv = cross(qv, v), uuv = cross(qv, uv);
    return v + ((uv * q.w) + uuv) * 2.0f;
}
inline float dot(const quat& a, const quat& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline quat conjugate(const quat& q) { return quat(q.w, -q.x, -q.y, -q.z); }
inline quat normalize(const quat& q)
{
    const float len = std::sqrt(dot(q, q));
    if (len <= 0.0f)
        return quat(1.0f, 0.0f, 0.0f, 0.0f);
    return q * (1.0f / len);
}
inline quat angleAxis(float angle, const vec3& axis)
{
    const float s = std::sin(angle * 0.5f);
    return quat(std::cos(angle * 0.5f), axis * s);
}
inline quat mix(const quat& a, const quat& b, float t)
{
    const float c = dot(a, b);
    if (c > 1.0f - 1e-6f)
        return quat(mix(a.w, b.w, t), mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t));
    const float ang = std::acos(c);
    return (a * std::sin((1.0f - t) * ang) + b * std::sin(t * ang)) * (1.0f / std::sin(ang));
}
inline quat slerp(const quat& a, const quat& b0, float t)
{
    quat b = b0;
    float c = dot(a, b);
    if (c < 0.0f)
    {
        b = -b0;
        c = -c;
    }
    if (c > 1.0f - 1e-6f)
        return quat(mix(a.w, b.w, t), mix(a.x, b.x, t), mix(a.y, b.y, t), mix(a.z, b.z, t));
    const float ang = std::acos(c);
    return (a * std::sin((1.0f - t) * ang) + b * std::sin(t * ang)) * (1.0f / std::sin(ang));
}
inline quat make_quat(const float* p) { return quat(p[3], p[0], p[1], p[2]); }

// 4x4 matrix, column-major: m[c] is column c
struct mat4
{
    vec4 col[4];
    mat4() : mat4(0.0f) {}
    explicit mat4(float s) { col[0] = vec4(s, 0, 0, 0); col[1] = vec4(0, s, 0, 0); col[2] = vec4(0, 0, s, 0); col[3] = vec4(0, 0, 0, s); }
    explicit mat4(double s) : mat4(float(s)) {}
    mat4(float x0, float y0, float z0, float w0, float x1, float y1, float z1, float w1, float x2, float y2, float z2, float w2, float x3, float y3,
         float z3, float w3)
    {
        col[0] = vec4(x0, y0, z0, w0);
        col[1] = vec4(x1, y1, z1, w1);
        col[2] = vec4(x2, y2, z2, w2);
        col[3] = vec4(x3, y3, z3, w3);
    }
    // rotation matrix of a quaternion (mat4_cast)
    explicit mat4(const quat& q)
    {
        const float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z, qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z, qwx = q.w * q.x,
                    qwy = q.w * q.y, qwz = q.w * q.z;
        col[0] = vec4(1.0f - 2.0f * (qyy + qzz), 2.0f * (qxy + qwz), 2.0f * (qxz - qwy), 0.0f);
        col[1] = vec4(2.0f * (qxy - qwz), 1.0f - 2.0f * (qxx + qzz), 2.0f * (qyz + qwx), 0.0f);
        col[2] = vec4(2.0f * (qxz + qwy), 2.0f * (qyz - qwx), 1.0f - 2.0f * (qxx + qyy), 0.0f);
        col[3] = vec4(0.0f, 0.0f, 0.0f, 1.0f);
    }
    vec4& operator[](int c) { return col[c]; }
    const vec4& operator[](int c) const { return col[c]; }
};
inline vec4 operator*(const mat4& m, const vec4& v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w; }
inline mat4 operator*(const mat4& a, const mat4& b)
{
    mat4 r(0.0f);
    for (int c = 0; c < 4; ++c)
        r[c] = a * b[c];
    return r;
}
inline bool operator==(const mat4& a, const mat4& b) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3]; }
inline bool operator!=(const mat4& a, const mat4& b) { return !(a == b); }
inline mat4 transpose(const mat4& m)
{
    mat4 r(0.0f);
    for (int c = 0; c < 4; ++c)
        for (int k = 0; k < 4; ++k)
            r[c][k] = m[k][c];
    return r;
}
inline mat4 translate(const mat4& m, const vec3& v)
{
    mat4 r = m;
    r[3] = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3];
    return r;
}
inline mat4 scale(const mat4& m, const vec3& v)
{
    mat4 r(0.0f);
    r[0] = m[0] * v.x;
    r[1] = m[1] * v.y;
    r[2] = m[2] * v.z;
    r[3] = m[3];
    return r;
}
inline mat4 mat4_cast(const quat& q) { return mat4(q); }
inline mat4 toMat4(const quat& q) { return mat4(q); }
inline mat4 inverse(const mat4& m)
{
    // cofactor expansion (double precision accumulate)
    double a[16], inv[16];
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r)
            a[c * 4 + r] = m[c][r];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    const double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    mat4 r(0.0f);
    for (int c = 0; c < 4; ++c)
        for (int k = 0; k < 4; ++k)
            r[c][k] = float(inv[c * 4 + k] / det);
    return r;
}
// right-handed, depth 0..1 (GLM_FORCE_DEPTH_ZERO_TO_ONE is defined by the reference's headers)
inline mat4 perspective(float fovy, float aspect, float zNear, float zFar)
{
    const float t = std::tan(fovy / 2.0f);
    mat4 r(0.0f);
    r[0][0] = 1.0f / (aspect * t);
    r[1][1] = 1.0f / t;
    r[2][2] = zFar / (zNear - zFar);
    r[2][3] = -1.0f;
    r[3][2] = -(zFar * zNear) / (zFar - zNear);
    return r;
}
inline mat4 lookAt(const vec3& eye, const vec3& center, const vec3& up)
{
    const vec3 f = normalize(center - eye), s = normalize(cross(f, up)), u = cross(s, f);
    mat4 r(1.0f);
    r[0][0] = s.x; r[1][0] = s.y; r[2][0] = s.z;
    r[0][1] = u.x; r[1][1] = u.y; r[2][1] = u.z;
    r[0][2] = -f.x; r[1][2] = -f.y; r[2][2] = -f.z;
    r[3][0] = -dot(s, eye); r[3][1] = -dot(u, eye); r[3][2] = dot(f, eye);
    return r;
}
inline float length2(const vec3& v) { return dot(v, v); }

inline const float* value_ptr(const mat4& m) { return &m.col[0].x; }
inline float* value_ptr(mat4& m) { return &m.col[0].x; }
inline const float* value_ptr(const vec3& v) { return &v.x; }
inline const float* value_ptr(const vec4& v) { return &v.x; }
inline const float* value_ptr(const quat& q) { return &q.x; }

// gtx/compatibility names
typedef vec2 float2;
typedef vec3 float3;
typedef vec4 float4;
typedef mat4 float4x4;
typedef mat4 mat4x4;
typedef vec2 fvec2;
typedef vec3 fvec3;
typedef vec4 fvec4;
struct ivec2 { int x = 0, y = 0; };
struct uvec2 { unsigned x = 0, y = 0; };

} // namespace glm
