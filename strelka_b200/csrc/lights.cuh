// Device light sampling / pdfs, following the reference's include/render/Lights.h expression by
// expression (uniform rect :277-289, spherical rectangle :86-189,245-275, sphere :335-362, distant
// cone :302-333, emitter-hit pdfs :201-243, balance heuristic :28-31).  Quirks kept: Q6 (whole-sphere
// constant pdf), Q7 (disc lights are never sampled), double-precision intermediates in SampleCone.
#pragma once
#include "hd.cuh"
#include "../../include/sb/sb_api.h"

namespace sb
{

struct LightSample // LightSampleData, Lights.h:16-26
{
    float3 pointOnLight;
    float pdf;
    float3 normal;
    float area;
    float3 L;
    float distToLight;
};

SB_HD float3 lpt(const sb_light& l, int i)
{
    return mk3(l.points[i][0], l.points[i][1], l.points[i][2]);
}
SB_HD float mis_balance(float a, float b)
{
    return 1.0f / (1.0f + (b / a));
}
SB_HD float light_area(const sb_light& l)
{
    float area = 0.0f;
    if (l.type == 0)
    {
        const float3 e1 = lpt(l, 1) - lpt(l, 0);
        const float3 e2 = lpt(l, 3) - lpt(l, 0);
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
SB_HD float3 light_normal(const sb_light& l, const float3& hitPoint)
{
    float3 n = mk3(0.0f);
    if (l.type == 0)
    {
        const float3 e1 = lpt(l, 1) - lpt(l, 0);
        const float3 e2 = lpt(l, 3) - lpt(l, 0);
        n = -normalize(cross(e1, e2));
    }
    else if (l.type == 1)
    {
        n = mk3(l.normal[0], l.normal[1], l.normal[2]);
    }
    else if (l.type == 2)
    {
        n = normalize(hitPoint - lpt(l, 1));
    }
    return n;
}
SB_HD void fill_light_data(const sb_light& l, const float3& hitPoint, LightSample& s)
{
    s.area = light_area(l);
    s.normal = light_normal(l, hitPoint);
    const float3 toLight = s.pointOnLight - hitPoint;
    const float len = length(toLight);
    s.L = toLight / len;
    s.distToLight = len;
}

struct SphQuad
{
    float3 o, x, y, z;
    float z0, z0sq, x0, y0, y0sq, x1, y1, y1sq, b0, b1, b0sq, k, S;
};
SB_HD SphQuad sphquad_init(const sb_light& l, const float3& o)
{
    SphQuad q;
    const float3 ex = lpt(l, 1) - lpt(l, 0);
    const float3 ey = lpt(l, 3) - lpt(l, 0);
    const float3 s = lpt(l, 0);
    const float exl = length(ex);
    const float eyl = length(ey);
    q.o = o;
    q.x = ex / exl;
    q.y = ey / eyl;
    q.z = cross(q.x, q.y);
    const float3 d = s - o;
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
    const float3 v00 = mk3(q.x0, q.y0, q.z0);
    const float3 v01 = mk3(q.x0, q.y1, q.z0);
    const float3 v10 = mk3(q.x1, q.y0, q.z0);
    const float3 v11 = mk3(q.x1, q.y1, q.z0);
    const float3 n0 = normalize(cross(v00, v10));
    const float3 n1 = normalize(cross(v10, v11));
    const float3 n2 = normalize(cross(v11, v01));
    const float3 n3 = normalize(cross(v01, v00));
    const float g0 = acosf(-dot(n0, n1));
    const float g1 = acosf(-dot(n1, n2));
    const float g2 = acosf(-dot(n2, n3));
    const float g3 = acosf(-dot(n3, n0));
    q.b0 = n0.z;
    q.b1 = n2.z;
    q.b0sq = q.b0 * q.b0;
    q.k = 2.0f * kPi - g2 - g3;
    q.S = g0 + g1 - q.k;
    return q;
}
SB_HD float3 sphquad_sample(const SphQuad& q, float u, float v)
{
    const float au = u * q.S + q.k;
    const float fu = (cosf(au) * q.b0 - q.b1) / sinf(au);
    float cu = 1.0f / sqrtf(fu * fu + q.b0sq) * (fu > 0.0f ? 1.0f : -1.0f);
    cu = clampf(cu, -1.0f, 1.0f);
    float xu = -(cu * q.z0) / sqrtf(1.0f - cu * cu);
    xu = clampf(xu, q.x0, q.x1);
    const float d = sqrtf(xu * xu + q.z0sq);
    const float h0 = q.y0 / sqrtf(d * d + q.y0sq);
    const float h1 = q.y1 / sqrtf(d * d + q.y1sq);
    const float hv = h0 + v * (h1 - h0);
    const float hv2 = hv * hv;
    const float eps = 1e-5f;
    const float yv = (hv < 1.0f - eps) ? (hv * d) / sqrtf(1 - hv2) : q.y1;
    return q.o + xu * q.x + yv * q.y + q.z0 * q.z;
}
SB_HD float rect_light_pdf(const sb_light& l, const float3& lightHit, const float3& surfaceHit)
{
    LightSample s;
    s.pointOnLight = lightHit;
    fill_light_data(l, surfaceHit, s);
    return s.distToLight * s.distToLight / (dot(-s.L, s.normal) * s.area);
}
// getLightPdf(l, lightHitPoint, surfaceHitPoint), Lights.h:221-243
template <bool RECT_UNIFORM = false>
SB_HD float light_pdf(const sb_light& l, const float3& lightHit, const float3& surfaceHit)
{
    if (RECT_UNIFORM)
        return rect_light_pdf(l, lightHit, surfaceHit);
    switch (l.type)
    {
    case 0:
        return rect_light_pdf(l, lightHit, surfaceHit);
    case 2:
        return 1.0f / (4.0f * kPi);
    case 3:
        return 1.0f / (2.0f * kPi * (1.0f - cosf(l.half_angle)));
    default:
        break;
    }
    return 0.0f;
}
SB_HD LightSample sample_rect_uniform(const sb_light& l, float u, float v, const float3& hitPoint)
{
    LightSample s;
    const float3 e1 = lpt(l, 1) - lpt(l, 0);
    const float3 e2 = lpt(l, 3) - lpt(l, 0);
    s.pointOnLight = lpt(l, 0) + e1 * u + e2 * v;
    fill_light_data(l, hitPoint, s);
    s.pdf = s.distToLight * s.distToLight / (-dot(s.L, s.normal) * s.area);
    return s;
}
SB_HD LightSample sample_rect_sphquad(const sb_light& l, float u, float v, const float3& hitPoint)
{
    LightSample s;
    const float3 e1 = lpt(l, 1) - lpt(l, 0);
    const float3 e2 = lpt(l, 3) - lpt(l, 0);
    const SphQuad q = sphquad_init(l, hitPoint);
    if (q.S <= 0.0f)
    {
        s.pdf = 0.0f;
        s.pointOnLight = lpt(l, 0) + e1 * u + e2 * v;
        fill_light_data(l, hitPoint, s);
        return s;
    }
    if (q.S < 1e-3f)
    {
        s.pointOnLight = lpt(l, 0) + e1 * u + e2 * v;
        fill_light_data(l, hitPoint, s);
        s.pdf = s.distToLight * s.distToLight / (-dot(s.L, s.normal) * s.area);
        return s;
    }
    s.pointOnLight = sphquad_sample(q, u, v);
    fill_light_data(l, hitPoint, s);
    s.pdf = 1.0f / q.S;
    return s;
}
SB_HD void coord_system(const float3& N, float3& Nt, float3& Nb)
{
    if (fabsf(N.x) > fabsf(N.y))
    {
        const float invLen = 1.0f / sqrtf(N.x * N.x + N.z * N.z);
        Nt = mk3(-N.z * invLen, 0.0f, N.x * invLen);
    }
    else
    {
        const float invLen = 1.0f / sqrtf(N.y * N.y + N.z * N.z);
        Nt = mk3(0.0f, N.z * invLen, -N.y * invLen);
    }
    Nb = cross(N, Nt);
}
// SampleCone, Lights.h:302-317: the reference's double literals promote these to double
SB_HD float3 sample_cone(float ux, float uy, float angle, const float3& direction, float& pdf)
{
    const float phi = float(2.0 * double(kPi) * double(ux));
    const float cosTheta = float(1.0 - double(uy) * (1.0 - double(cosf(angle))));
    const float sinTheta = float(sqrt(1.0 - double(cosTheta * cosTheta)));
    float3 u, v;
    coord_system(direction, u, v);
    const float3 dir = normalize(cosf(phi) * sinTheta * u + sinf(phi) * sinTheta * v + cosTheta * direction);
    pdf = float(1.0 / (2.0 * double(kPi) * (1.0 - double(cosf(angle)))));
    return dir;
}
SB_HD LightSample sample_distant(const sb_light& l, float u, float v)
{
    LightSample s;
    float pdf = 0.0f;
    const float3 n = mk3(l.normal[0], l.normal[1], l.normal[2]);
    const float3 c = sample_cone(u, v, l.half_angle, -n, pdf);
    s.area = 0.0f;
    s.distToLight = 1e9f;
    s.L = c;
    s.normal = n;
    s.pdf = pdf;
    s.pointOnLight = c;
    return s;
}
SB_HD LightSample sample_sphere(const sb_light& l, float u, float v, const float3& hitPoint)
{
    LightSample s;
    const float cosTheta = 1.0f - 2.0f * u;
    const float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    const float phi = 2.0f * kPi * v;
    const float radius = l.points[0][0];
    const float3 dir = mk3(sinTheta * cosf(phi), sinTheta * sinf(phi), cosTheta);
    const float3 lightPoint = lpt(l, 1) + radius * dir;
    s.L = normalize(lightPoint - hitPoint);
    s.distToLight = length(lightPoint - hitPoint);
    s.area = 0.0f;
    s.normal = dir;
    s.pdf = 1.0f / (4.0f * kPi);
    s.pointOnLight = lightPoint;
    return s;
}
// the switch of sampleLight(), closest_hit.cu:266-291
// RECT_UNIFORM = true: the caller guarantees rect lights sampled uniformly only (compiles the other samplers out)
template <bool RECT_UNIFORM = false>
SB_HD LightSample sample_light(const sb_light& l, float u, float v, const float3& hitPoint, uint32_t rectMethod)
{
    if (RECT_UNIFORM)
        return sample_rect_uniform(l, u, v, hitPoint);
    LightSample s;
    s.pointOnLight = mk3(0.0f);
    s.pdf = 0.0f;
    s.normal = mk3(0.0f);
    s.area = 0.0f;
    s.L = mk3(0.0f);
    s.distToLight = 0.0f;
    switch (l.type)
    {
    case 0:
        s = (rectMethod == 0) ? sample_rect_uniform(l, u, v, hitPoint) : sample_rect_sphquad(l, u, v, hitPoint);
        break;
    case 2:
        s = sample_sphere(l, u, v, hitPoint);
        break;
    case 3:
        s = sample_distant(l, u, v);
        break;
    default:
        break;
    }
    return s;
}

} // namespace sb
