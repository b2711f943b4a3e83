// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the reference light sampling (include/render/Lights.h): uniform and
// spherical-rectangle (Urena et al. 2013) rect sampling, whole-sphere sampling, distant cone
// sampling, the pdfs used on emitter hits and the balance heuristic.  Parity target: <= 1 ulp
// of the reference header compiled on the host (oracle/ref_crosscheck.cpp), frozen in
// tests/golden/lights.json.
#pragma once
#include "vec.h"
#include "../include/sb/sb_api.h"

namespace orc
{

static constexpr float kPi = 3.14159265358979323846f; // M_PIf

// LightSampleData, Lights.h:16-26
struct LightSample
{
    f3 pointOnLight;
    float pdf;
    f3 normal;
    float area;
    f3 L;
    float distToLight;
};

inline f3 lp(const sb_light& l, int i)
{
    return f3{ l.points[i][0], l.points[i][1], l.points[i][2] };
}

// misWeightBalance(), Lights.h:28-31
inline float mis_balance(float a, float b)
{
    return 1.0f / (1.0f + (b / a));
}

// calcLightArea(), Lights.h:33-52
inline float light_area(const sb_light& l)
{
    float area = 0.0f;
    if (l.type == 0)
    {
        const f3 e1 = lp(l, 1) - lp(l, 0);
        const f3 e2 = lp(l, 3) - lp(l, 0);
        area = length(cross(e1, e2));
    }
    else if (l.type == 1)
    {
        area = kPi * l.points[0][0] * l.points[0][0];
    }
    else if (l.type == 2)
    {
        area = 4.0f * kPi * l.points[0][0] * l.points[0][0];
    }
    return area;
}

// calcLightNormal(), Lights.h:54-74
inline f3 light_normal(const sb_light& l, const f3& hitPoint)
{
    f3 n = mk3(0.0f);
    if (l.type == 0)
    {
        const f3 e1 = lp(l, 1) - lp(l, 0);
        const f3 e2 = lp(l, 3) - lp(l, 0);
        n = -normalize(cross(e1, e2));
    }
    else if (l.type == 1)
    {
        n = f3{ l.normal[0], l.normal[1], l.normal[2] };
    }
    else if (l.type == 2)
    {
        n = normalize(hitPoint - lp(l, 1));
    }
    return n;
}

// fillLightData(), Lights.h:76-84
inline void fill_light_data(const sb_light& l, const f3& hitPoint, LightSample& s)
{
    s.area = light_area(l);
    s.normal = light_normal(l, hitPoint);
    const f3 toLight = s.pointOnLight - hitPoint;
    const float len = length(toLight);
    s.L = toLight / len;
    s.distToLight = len;
}

// SphQuad + init(), Lights.h:86-161
struct SphQuad
{
    f3 o, x, y, z;
    float z0, z0sq;
    float x0, y0, y0sq;
    float x1, y1, y1sq;
    float b0, b1, b0sq, k;
    float S;
};

inline SphQuad sphquad_init(const sb_light& l, const f3& o)
{
    SphQuad q;
    const f3 ex = lp(l, 1) - lp(l, 0);
    const f3 ey = lp(l, 3) - lp(l, 0);
    const f3 s = lp(l, 0);
    const float exl = length(ex);
    const float eyl = length(ey);
    q.o = o;
    q.x = ex / exl;
    q.y = ey / eyl;
    q.z = cross(q.x, q.y);
    const f3 d = s - o;
    q.z0 = dot(d, q.z);
    if (q.z0 > 0)
    {
        q.z *= -1.0f;
        q.z0 *= -1.0f;
    }
    q.z0sq = q.z0 * q.z0;
    q.x0 = dot(d, q.x);
    q.y0 = dot(d, q.y);
    q.x1 = q.x0 + exl;
    q.y1 = q.y0 + eyl;
    q.y0sq = q.y0 * q.y0;
    q.y1sq = q.y1 * q.y1;
    const f3 v00{ q.x0, q.y0, q.z0 };
    const f3 v01{ q.x0, q.y1, q.z0 };
    const f3 v10{ q.x1, q.y0, q.z0 };
    const f3 v11{ q.x1, q.y1, q.z0 };
    const f3 n0 = normalize(cross(v00, v10));
    const f3 n1 = normalize(cross(v10, v11));
    const f3 n2 = normalize(cross(v11, v01));
    const f3 n3 = normalize(cross(v01, v00));
    const float g0 = std::acos(-dot(n0, n1));
    const float g1 = std::acos(-dot(n1, n2));
    const float g2 = std::acos(-dot(n2, n3));
    const float g3 = std::acos(-dot(n3, n0));
    q.b0 = n0.z;
    q.b1 = n2.z;
    q.b0sq = q.b0 * q.b0;
    q.k = 2.0f * kPi - g2 - g3;
    q.S = g0 + g1 - q.k;
    return q;
}

// SphQuadSample(), Lights.h:163-189
inline f3 sphquad_sample(const SphQuad& q, float u, float v)
{
    const float au = u * q.S + q.k;
    const float fu = (std::cos(au) * q.b0 - q.b1) / std::sin(au);
    float cu = 1.0f / std::sqrt(fu * fu + q.b0sq) * (fu > 0.0f ? 1.0f : -1.0f);
    cu = clampf(cu, -1.0f, 1.0f);
    float xu = -(cu * q.z0) / std::sqrt(1.0f - cu * cu);
    xu = clampf(xu, q.x0, q.x1);
    const float d = std::sqrt(xu * xu + q.z0sq);
    const float h0 = q.y0 / std::sqrt(d * d + q.y0sq);
    const float h1 = q.y1 / std::sqrt(d * d + q.y1sq);
    const float hv = h0 + v * (h1 - h0);
    const float hv2 = hv * hv;
    const float eps = 1e-5f;
    const float yv = (hv < 1.0f - eps) ? (hv * d) / std::sqrt(1 - hv2) : q.y1;
    return q.o + xu * q.x + yv * q.y + q.z0 * q.z;
}

// getRectLightPdf(), Lights.h:201-209
inline float rect_light_pdf(const sb_light& l, const f3& lightHit, const f3& surfaceHit)
{
    LightSample s{};
    s.pointOnLight = lightHit;
    fill_light_data(l, surfaceHit, s);
    return s.distToLight * s.distToLight / (dot(-s.L, s.normal) * s.area);
}

// getLightPdf(light, lightHitPoint, surfaceHitPoint), Lights.h:221-243 (+211-219)
inline float light_pdf(const sb_light& l, const f3& lightHit, const f3& surfaceHit)
{
    switch (l.type)
    {
    case 0:
        return rect_light_pdf(l, lightHit, surfaceHit);
    case 2:
        return 1.0f / (4.0f * kPi);
    case 3:
        // getDirectLightPdf: the reference calls double cos() on a float and rounds the
        // quotient back to float (Lights.h:211-214)
        return 1.0f / (2.0f * kPi * (1.0f - std::cos(l.half_angle)));
    default:
        break;
    }
    return 0.0f;
}

// SampleRectLightUniform(), Lights.h:277-289
inline LightSample sample_rect_uniform(const sb_light& l, float u, float v, const f3& hitPoint)
{
    LightSample s;
    const f3 e1 = lp(l, 1) - lp(l, 0);
    const f3 e2 = lp(l, 3) - lp(l, 0);
    s.pointOnLight = lp(l, 0) + e1 * u + e2 * v;
    fill_light_data(l, hitPoint, s);
    s.pdf = s.distToLight * s.distToLight / (-dot(s.L, s.normal) * s.area);
    return s;
}

// SampleRectLight(), Lights.h:245-275
inline LightSample sample_rect_sphquad(const sb_light& l, float u, float v, const f3& hitPoint)
{
    LightSample s;
    const f3 e1 = lp(l, 1) - lp(l, 0);
    const f3 e2 = lp(l, 3) - lp(l, 0);
    const SphQuad q = sphquad_init(l, hitPoint);
    if (q.S <= 0.0f)
    {
        s.pdf = 0.0f;
        s.pointOnLight = lp(l, 0) + e1 * u + e2 * v;
        fill_light_data(l, hitPoint, s);
        return s;
    }
    if (q.S < 1e-3f)
    {
        s.pointOnLight = lp(l, 0) + e1 * u + e2 * v;
        fill_light_data(l, hitPoint, s);
        s.pdf = s.distToLight * s.distToLight / (-dot(s.L, s.normal) * s.area);
        return s;
    }
    s.pointOnLight = sphquad_sample(q, u, v);
    fill_light_data(l, hitPoint, s);
    s.pdf = 1.0f / q.S;
    return s;
}

// createCoordinateSystem(), Lights.h:291-300
inline void coord_system(const f3& N, f3& Nt, f3& Nb)
{
    if (std::fabs(N.x) > std::fabs(N.y))
    {
        const float invLen = 1.0f / std::sqrt(N.x * N.x + N.z * N.z);
        Nt = f3{ -N.z * invLen, 0.0f, N.x * invLen };
    }
    else
    {
        const float invLen = 1.0f / std::sqrt(N.y * N.y + N.z * N.z);
        Nt = f3{ 0.0f, N.z * invLen, -N.y * invLen };
    }
    Nb = cross(N, Nt);
}

// SampleCone(), Lights.h:302-317.  The reference mixes double literals (2.0, 1.0) with floats,
// so phi/cosTheta/sinTheta/pdf are computed in double and rounded to float on assignment.
inline f3 sample_cone(float ux, float uy, float angle, const f3& direction, float& pdf)
{
    const float phi = float(2.0 * double(kPi) * double(ux));
    const float cosTheta = float(1.0 - double(uy) * (1.0 - double(std::cos(angle))));
    const float sinTheta = float(std::sqrt(1.0 - double(cosTheta * cosTheta)));
    f3 u, v;
    coord_system(direction, u, v);
    const f3 dir = normalize(std::cos(phi) * sinTheta * u + std::sin(phi) * sinTheta * v + cosTheta * direction);
    pdf = float(1.0 / (2.0 * double(kPi) * (1.0 - double(std::cos(angle)))));
    return dir;
}

// SampleDistantLight(), Lights.h:319-333
inline LightSample sample_distant(const sb_light& l, float u, float v, const f3&)
{
    LightSample s;
    float pdf = 0.0f;
    const f3 n{ l.normal[0], l.normal[1], l.normal[2] };
    const f3 c = sample_cone(u, v, l.half_angle, -n, pdf);
    s.area = 0.0f;
    s.distToLight = 1e9f;
    s.L = c;
    s.normal = n;
    s.pdf = pdf;
    s.pointOnLight = c;
    return s;
}

// SampleSphereLight(), Lights.h:335-362 (quirk Q6: whole sphere, constant pdf)
inline LightSample sample_sphere(const sb_light& l, float u, float v, const f3& hitPoint)
{
    LightSample s;
    const float cosTheta = 1.0f - 2.0f * u;
    const float sinTheta = std::sqrt(1.0f - cosTheta * cosTheta);
    const float phi = 2.0f * kPi * v;
    const float radius = l.points[0][0];
    const f3 dir{ sinTheta * std::cos(phi), sinTheta * std::sin(phi), cosTheta };
    const f3 lightPoint = lp(l, 1) + radius * dir;
    s.L = normalize(lightPoint - hitPoint);
    s.distToLight = length(lightPoint - hitPoint);
    s.area = 0.0f;
    s.normal = dir;
    s.pdf = 1.0f / (4.0f * kPi);
    s.pointOnLight = lightPoint;
    return s;
}

// The switch of sampleLight(), closest_hit.cu:266-291.  Disc lights (type 1) are never
// sampled (quirk Q7): the sample stays zero-initialised.
inline LightSample sample_light(const sb_light& l, float u, float v, const f3& hitPoint, uint32_t rectMethod)
{
    LightSample s{};
    switch (l.type)
    {
    case 0:
        s = (rectMethod == 0) ? sample_rect_uniform(l, u, v, hitPoint) : sample_rect_sphquad(l, u, v, hitPoint);
        break;
    case 2:
        s = sample_sphere(l, u, v, hitPoint);
        break;
    case 3:
        s = sample_distant(l, u, v, hitPoint);
        break;
    default:
        break;
    }
    return s;
}

} // namespace orc
