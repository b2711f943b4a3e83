// Hair fibre BSDF for SB_MATERIAL_HAIR: the near-field model of Chiang, Bitterli, Tappan and Burley, "A Practical and
// Controllable Hair and Fur Model for Production Path Tracing" (2016) -- the model behind MDL's df::chiang_hair_bsdf,
// which is what the reference's hair materials compile to (mdlPtxCodeGen.cpp:143-155; the .mdl asset itself is not in
// the tree: PARITY UNPINNED).  Lobes R, TT, TRT plus the residual; longitudinal scattering M_p
// (d'Eon's energy-conserving Gaussian), azimuthal scattering N_p (trimmed logistic), attenuations A_p from Fresnel and
// absorption.  The reference feeds every curve hit the texture coordinate (0.5, 0.5, 0.5) (closest_hit.cu:446), so the
// azimuthal offset h = 2 * 0.5 - 1 is 0 for every hit; the fibre frame is (curveTangent, surfaceNormal)
// (closest_hit.cu:435-440).
//
// Protocol (closest_hit.cu:474-605): evaluate returns the scattering function times the cosine (all of it in the
// glossy slot), sample returns k2, bsdf_over_pdf, pdf and GLOSSY | REFLECTION or GLOSSY | TRANSMISSION, the latter
// when k2 leaves through the other side of the geometric surface (so `inside` toggles and the continuation ray starts
// below the surface, closest_hit.cu:591-600).
#pragma once
#include "hd.cuh"
#include "../../include/sb/sb_api.h"

namespace sb
{

constexpr int kHairLobes = 3; // R, TT, TRT; index kHairLobes is the residual
constexpr float kHairH = 0.0f; // azimuthal offset (fixed by the reference's texture coordinate)

struct HairLobes
{
    float v[kHairLobes + 1]; // longitudinal variances
    float s; // logistic scale of the azimuthal lobes
    float sinA[3], cosA[3]; // sin / cos of alpha, 2 alpha, 4 alpha (cuticle tilt)
    float eta;
    float3 sigmaA;
};

SB_HD HairLobes hair_init(const sb_material& m)
{
    HairLobes L;
    const float bm = clampf(m.hair_roughness_lon, 0.01f, 1.0f), bn = clampf(m.hair_roughness_azi, 0.01f, 1.0f);
    const float bm2 = bm * bm, bm4 = bm2 * bm2, bm8 = bm4 * bm4, bm16 = bm8 * bm8;
    const float r = 0.726f * bm + 0.812f * bm2 + 3.7f * (bm16 * bm4);
    L.v[0] = r * r;
    L.v[1] = 0.25f * L.v[0];
    L.v[2] = 4.0f * L.v[0];
    L.v[3] = L.v[2];
    const float bn2 = bn * bn, bn4 = bn2 * bn2, bn8 = bn4 * bn4, bn16 = bn8 * bn8;
    L.s = 0.626657069f * (0.265f * bn + 1.194f * bn2 + 5.372f * (bn16 * bn4 * bn2)); // sqrt(pi / 8) * (...)
    L.sinA[0] = sinf(m.hair_cuticle_angle);
    L.cosA[0] = sqrtf(fmaxf(0.0f, 1.0f - L.sinA[0] * L.sinA[0]));
    for (int i = 1; i < 3; ++i)
    {
        L.sinA[i] = 2.0f * L.cosA[i - 1] * L.sinA[i - 1];
        L.cosA[i] = L.cosA[i - 1] * L.cosA[i - 1] - L.sinA[i - 1] * L.sinA[i - 1];
    }
    L.eta = m.ior > 1.0f ? m.ior : 1.55f;
    L.sigmaA = mk3(fmaxf(m.hair_absorption[0], 0.0f), fmaxf(m.hair_absorption[1], 0.0f), fmaxf(m.hair_absorption[2], 0.0f));
    return L;
}

// log of the modified Bessel function of the first kind, order 0: power series up to x = 8 (16 terms: truncation
// below 1e-7 relative), four terms of the asymptotic expansion e^x / sqrt(2 pi x) * (1 + 1/(8x) + 9/(128x^2) + ...)
// beyond (below 4e-5 relative at x = 8, falling with x^-4)
SB_HD float hair_log_i0(float x)
{
    if (x > 8.0f)
    {
        const float r = 1.0f / x;
        const float tail = 1.0f + r * (0.125f + r * (0.0703125f + r * 0.0732421875f));
        return x - 0.5f * logf(2.0f * kPi * x) + logf(tail);
    }
    float sum = 0.0f, term = 1.0f; // term_i = (x/2)^(2i) / (i!)^2
    const float q = 0.25f * x * x;
    for (int i = 0; i < 16; ++i)
    {
        sum += term;
        term = term * q / (float(i + 1) * float(i + 1));
    }
    return logf(sum);
}
// longitudinal scattering function M_p (d'Eon et al. 2011): exp(-b) I0(a) / (2 v sinh(1/v)), a = cos cos / v,
// b = sin sin / v, evaluated in the log domain so that hair-smooth lobes (v ~ 1e-5) neither overflow nor cancel
SB_HD float hair_mp(float cosI, float cosO, float sinI, float sinO, float v)
{
    const float a = cosI * cosO / v, b = sinI * sinO / v, iv = 1.0f / v;
    return expf(hair_log_i0(a) - b - iv) * iv / (1.0f - expf(-2.0f * iv));
}
// unpolarised Fresnel reflectance of a dielectric of relative index eta, seen from outside under cosine c
SB_HD float hair_fresnel(float c, float eta)
{
    c = clampf(c, 0.0f, 1.0f);
    const float sinT2 = (1.0f - c * c) / (eta * eta);
    if (sinT2 >= 1.0f)
        return 1.0f;
    const float cosT = sqrtf(1.0f - sinT2);
    const float rpar = (eta * c - cosT) / (eta * c + cosT), rper = (c - eta * cosT) / (c + eta * cosT);
    return 0.5f * (rpar * rpar + rper * rper);
}
// attenuations A_0..A_3 (A_3 = all higher orders)
SB_HD void hair_ap(float cosThetaO, float eta, float h, const float3& T, float3 ap[kHairLobes + 1])
{
    const float cosGammaO = sqrtf(fmaxf(0.0f, 1.0f - h * h));
    const float f = hair_fresnel(cosThetaO * cosGammaO, eta);
    ap[0] = mk3(f);
    ap[1] = T * ((1.0f - f) * (1.0f - f));
    ap[2] = ap[1] * T * f;
    const float3 den = mk3(1.0f) - T * f;
    ap[3] = ap[2] * T * f / den;
}
SB_HD float hair_logistic(float x, float s)
{
    x = fabsf(x);
    const float e = expf(-x / s);
    return e / (s * (1.0f + e) * (1.0f + e));
}
SB_HD float hair_logistic_cdf(float x, float s)
{
    return 1.0f / (1.0f + expf(-x / s));
}
SB_HD float hair_trimmed_logistic(float x, float s)
{
    return hair_logistic(x, s) / (hair_logistic_cdf(kPi, s) - hair_logistic_cdf(-kPi, s));
}
SB_HD float hair_sample_trimmed_logistic(float u, float s)
{
    const float lo = hair_logistic_cdf(-kPi, s), k = hair_logistic_cdf(kPi, s) - lo;
    const float x = -s * logf(1.0f / (u * k + lo) - 1.0f);
    return clampf(x, -kPi, kPi);
}
// azimuthal scattering function N_p around the lobe centre Phi(p) = 2 p gammaT - 2 gammaO + p pi
SB_HD float hair_np(float phi, int p, float s, float gammaO, float gammaT)
{
    float d = phi - (2.0f * float(p) * gammaT - 2.0f * gammaO + float(p) * kPi);
    while (d > kPi)
        d -= 2.0f * kPi;
    while (d < -kPi)
        d += 2.0f * kPi;
    return hair_trimmed_logistic(d, s);
}
// outgoing-side angles after the cuticle tilt of lobe p
SB_HD void hair_tilt(const HairLobes& L, int p, float sinO, float cosO, float& sinOp, float& cosOp)
{
    if (p == 0)
    {
        sinOp = sinO * L.cosA[1] - cosO * L.sinA[1];
        cosOp = cosO * L.cosA[1] + sinO * L.sinA[1];
    }
    else if (p == 1)
    {
        sinOp = sinO * L.cosA[0] + cosO * L.sinA[0];
        cosOp = cosO * L.cosA[0] - sinO * L.sinA[0];
    }
    else
    {
        sinOp = sinO * L.cosA[2] + cosO * L.sinA[2];
        cosOp = cosO * L.cosA[2] - sinO * L.sinA[2];
    }
    cosOp = fabsf(cosOp);
}

// quantities that depend on the outgoing direction only
struct HairOut
{
    float sinO, cosO, phiO, gammaT;
    float3 ap[kHairLobes + 1];
    float apPdf[kHairLobes + 1];
};
SB_HD HairOut hair_outgoing(const HairLobes& L, const float3& wo)
{
    HairOut o;
    o.sinO = clampf(wo.x, -1.0f, 1.0f);
    o.cosO = sqrtf(fmaxf(0.0f, 1.0f - o.sinO * o.sinO));
    o.phiO = atan2f(wo.z, wo.y);
    // refraction into the fibre: modified index eta', transmittance along one internal chord
    const float sinT = o.sinO / L.eta;
    const float cosT = sqrtf(fmaxf(0.0f, 1.0f - sinT * sinT));
    const float etap = sqrtf(fmaxf(L.eta * L.eta - o.sinO * o.sinO, 0.0f)) / fmaxf(o.cosO, 1e-6f);
    const float sinGammaT = clampf(kHairH / etap, -1.0f, 1.0f);
    const float cosGammaT = sqrtf(fmaxf(0.0f, 1.0f - sinGammaT * sinGammaT));
    o.gammaT = asinf(sinGammaT);
    const float len = 2.0f * cosGammaT / fmaxf(cosT, 1e-6f);
    const float3 T = mk3(expf(-L.sigmaA.x * len), expf(-L.sigmaA.y * len), expf(-L.sigmaA.z * len));
    hair_ap(o.cosO, L.eta, kHairH, T, o.ap);
    float sum = 0.0f;
    for (int p = 0; p <= kHairLobes; ++p)
    {
        o.apPdf[p] = dot(o.ap[p], mk3(0.299f, 0.587f, 0.114f));
        sum += o.apPdf[p];
    }
    const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
    for (int p = 0; p <= kHairLobes; ++p)
        o.apPdf[p] *= inv;
    return o;
}

// scattering function times |cos| and the sampling density for local directions wo, wi (x along the fibre)
SB_HD void hair_eval_local(const HairLobes& L, const HairOut& o, const float3& wi, float3& fcos, float& pdf)
{
    const float sinI = clampf(wi.x, -1.0f, 1.0f);
    const float cosI = sqrtf(fmaxf(0.0f, 1.0f - sinI * sinI));
    const float phi = atan2f(wi.z, wi.y) - o.phiO;
    const float gammaO = asinf(kHairH);
    fcos = mk3(0.0f);
    pdf = 0.0f;
    for (int p = 0; p < kHairLobes; ++p)
    {
        float sinOp, cosOp;
        hair_tilt(L, p, o.sinO, o.cosO, sinOp, cosOp);
        const float mn = hair_mp(cosI, cosOp, sinI, sinOp, L.v[p]) * hair_np(phi, p, L.s, gammaO, o.gammaT);
        fcos += o.ap[p] * mn;
        pdf += o.apPdf[p] * mn;
    }
    const float mr = hair_mp(cosI, o.cosO, sinI, o.sinO, L.v[kHairLobes]) * (1.0f / (2.0f * kPi));
    fcos += o.ap[kHairLobes] * mr;
    pdf += o.apPdf[kHairLobes] * mr;
}

struct HairFrame
{
    float3 t, n, b; // fibre tangent, normal, binormal
};
SB_HD HairFrame hair_frame(const float3& tangent, const float3& normal)
{
    HairFrame f;
    f.t = normalize(tangent);
    float3 n = normal - f.t * dot(normal, f.t);
    const float l2 = dot(n, n);
    if (!(l2 > 1e-12f))
    {
        // degenerate (normal along the fibre): any perpendicular will do
        const float3 a = fabsf(f.t.x) < 0.9f ? mk3(1.0f, 0.0f, 0.0f) : mk3(0.0f, 1.0f, 0.0f);
        n = a - f.t * dot(a, f.t);
    }
    f.n = normalize(n);
    f.b = cross(f.t, f.n);
    return f;
}
SB_HD float3 hair_to_local(const HairFrame& f, const float3& w)
{
    return mk3(dot(w, f.t), dot(w, f.n), dot(w, f.b));
}

// Everything that depends on the material, the fibre frame and the outgoing direction k1 only: computed once per hit and
// shared by the BSDF sample and the NEE evaluation of the same bounce (closest_hit.cu:521 and :571 see the same state).
struct HairCtx
{
    HairLobes L;
    HairFrame F;
    HairOut o;
};
SB_HD HairCtx hair_prepare(const sb_material& m, const float3& normal, const float3& tangent, const float3& k1)
{
    HairCtx c;
    c.L = hair_init(m);
    c.F = hair_frame(tangent, normal);
    c.o = hair_outgoing(c.L, hair_to_local(c.F, k1));
    return c;
}

// evaluate: fcos = f * |cos| for the world direction k2 (towards the light)
SB_HD void hair_evaluate(const HairCtx& c, const float3& k2, float3& fcos, float& pdf)
{
    hair_eval_local(c.L, c.o, hair_to_local(c.F, k2), fcos, pdf);
}

// sample: xi.z picks the lobe, (xi.x, xi.y) the longitudinal angle, xi.w the azimuth.  Returns false for an absorbed sample.
SB_HD bool hair_sample(const HairCtx& c, const float4& xi, float3& k2, float3& weight, float& pdf)
{
    const HairLobes& L = c.L;
    const HairFrame& F = c.F;
    const HairOut& o = c.o;
    int p = 0;
    float u = xi.z;
    for (; p < kHairLobes; ++p)
    {
        if (u < o.apPdf[p])
            break;
        u -= o.apPdf[p];
    }
    float sinOp = o.sinO, cosOp = o.cosO;
    if (p < kHairLobes)
        hair_tilt(L, p, o.sinO, o.cosO, sinOp, cosOp);
    // longitudinal: M_p is sampled exactly (inverse of the exponential of the cosine)
    const float u1 = fmaxf(xi.x, 1e-5f);
    const float v = L.v[p];
    const float cosTheta = 1.0f + v * logf(u1 + (1.0f - u1) * expf(-2.0f / v));
    const float sinTheta = sqrtf(fmaxf(0.0f, 1.0f - cosTheta * cosTheta));
    float sphi, cphi;
    sincos2pi(xi.y, sphi, cphi);
    const float sinI = clampf(-cosTheta * sinOp + sinTheta * cphi * cosOp, -1.0f, 1.0f);
    const float cosI = sqrtf(fmaxf(0.0f, 1.0f - sinI * sinI));
    // azimuthal: logistic around the lobe centre, uniform for the residual
    const float gammaO = asinf(kHairH);
    const float dphi = (p < kHairLobes) ? (2.0f * float(p) * o.gammaT - 2.0f * gammaO + float(p) * kPi) + hair_sample_trimmed_logistic(xi.w, L.s)
                                        : 2.0f * kPi * xi.w;
    const float phiI = o.phiO + dphi;
    const float3 wi = mk3(sinI, cosI * cosf(phiI), cosI * sinf(phiI));
    float3 fcos;
    hair_eval_local(L, o, wi, fcos, pdf);
    if (!(pdf > 0.0f))
        return false;
    k2 = normalize(wi.x * F.t + wi.y * F.n + wi.z * F.b);
    weight = fcos / pdf;
    return true;
}

// one-shot forms (test hook, host emulation)
SB_HD void hair_evaluate(const sb_material& m, const float3& normal, const float3& tangent, const float3& k1, const float3& k2, float3& fcos,
                         float& pdf)
{
    const HairCtx c = hair_prepare(m, normal, tangent, k1);
    hair_evaluate(c, k2, fcos, pdf);
}
SB_HD bool hair_sample(const sb_material& m, const float3& normal, const float3& tangent, const float3& k1, const float4& xi, float3& k2,
                       float3& weight, float& pdf)
{
    const HairCtx c = hair_prepare(m, normal, tangent, k1);
    return hair_sample(c, xi, k2, weight, pdf);
}

} // namespace sb
