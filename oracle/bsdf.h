// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Closed-form BSDFs.  PARITY UNPINNED: the reference evaluates every material through NVIDIA's
// closed MDL SDK (mdlcode_init/sample/evaluate, closest_hit.cu:502,521,571) from .mdl sources
// that are not in its tree.  What IS pinned and reproduced here is the call protocol of
// closest_hit.cu:474-605: inputs (k1 = -ray_dir, xi = 4 randoms, shading normal, geometric
// normal), outputs (k2, bsdf_over_pdf, pdf, event_type bit mask; evaluate() returns
// bsdf_diffuse / bsdf_glossy WITH the cosine included, and a pdf).  The model definitions are
// this repository's (DESIGN.md "Materials"); the in-repo precedent for the
// diffuse model is the Metal backend's Lambert (src/render/metal/shaders/pathtrace.metal:181-201).
#pragma once
#include "vec.h"
#include "../include/sb/sb_api.h"
#include "hair.h"
#include "lights.h" // kPi
#include "../include/sb/sb_api.h"

namespace orc
{

// mi::neuraylib::Bsdf_event_type bits as used by closest_hit.cu:523-547,593
enum : int
{
    EV_ABSORB = 0,
    EV_DIFFUSE = 1,
    EV_GLOSSY = 2,
    EV_SPECULAR = 4,
    EV_REFLECTION = 8,
    EV_TRANSMISSION = 16
};

struct BsdfSample
{
    f3 k2;
    f3 bsdf_over_pdf;
    float pdf;
    int event;
};
struct BsdfEval
{
    f3 diffuse; // cos included
    f3 glossy; // cos included
    float pdf;
};

// Branch-light orthonormal basis around a unit vector (Duff et al. 2017)
inline void onb(const f3& n, f3& t, f3& b)
{
    const float sign = std::copysign(1.0f, n.z);
    const float a = -1.0f / (sign + n.z);
    const float c = n.x * n.y * a;
    t = f3{ 1.0f + sign * n.x * n.x * a, sign * c, -sign * n.x };
    b = f3{ c, sign + n.y * n.y * a, -n.y };
}

inline f3 cosine_hemisphere(float u1, float u2, const f3& n)
{
    const float r = std::sqrt(u1);
    float sphi, cphi;
    sincos2pi(u2, sphi, cphi);
    const float x = r * cphi;
    const float y = r * sphi;
    const float z = std::sqrt(std::fmax(0.0f, 1.0f - u1));
    f3 t, b;
    onb(n, t, b);
    return normalize(x * t + y * b + z * n);
}

inline float luminance(const f3& c)
{
    return dot(c, f3{ 0.299f, 0.587f, 0.114f });
}
inline float pow5(float x)
{
    const float x2 = x * x;
    return x2 * x2 * x;
}
inline float schlick(float f0, float c)
{
    return f0 + (1.0f - f0) * pow5(1.0f - c);
}
inline f3 schlick3(const f3& f0, float c)
{
    const float w = pow5(1.0f - c);
    return f0 + (mk3(1.0f) - f0) * w;
}
// see the device file for why sin^2 comes from the cross product
inline float ggx_d(float a, float nh, float sin2h)
{
    const float a2 = a * a;
    const float d = a2 * nh * nh + sin2h;
    return a2 / (kPi * d * d);
}
inline float smith_g1(float a, float nx)
{
    const float a2 = a * a;
    return 2.0f * nx / (nx + std::sqrt(a2 + (1.0f - a2) * nx * nx));
}

// Heitz 2018 visible-normal sampling; v and result in world space around n
inline f3 ggx_sample_vndf(float a, const f3& n, const f3& v, float u1, float u2)
{
    f3 t, b;
    onb(n, t, b);
    const f3 vl{ dot(v, t), dot(v, b), dot(v, n) };
    const f3 vh = normalize(f3{ a * vl.x, a * vl.y, vl.z });
    const float lensq = vh.x * vh.x + vh.y * vh.y;
    const f3 T1 = lensq > 0.0f ? f3{ -vh.y, vh.x, 0.0f } * (1.0f / std::sqrt(lensq)) : f3{ 1.0f, 0.0f, 0.0f };
    const f3 T2 = cross(vh, T1);
    const float r = std::sqrt(u1);
    float sphi, cphi;
    sincos2pi(u2, sphi, cphi);
    const float t1 = r * cphi;
    float t2 = r * sphi;
    const float s = 0.5f * (1.0f + vh.z);
    t2 = (1.0f - s) * std::sqrt(std::fmax(0.0f, 1.0f - t1 * t1)) + s * t2;
    const f3 nh = t1 * T1 + t2 * T2 + std::sqrt(std::fmax(0.0f, 1.0f - t1 * t1 - t2 * t2)) * vh;
    const f3 hl = normalize(f3{ a * nh.x, a * nh.y, std::fmax(0.0f, nh.z) });
    return normalize(hl.x * t + hl.y * b + hl.z * n);
}

// ---- UsdPreviewSurface lobes ------------------------------------------------------------------
struct UpsLobes
{
    f3 diffAlbedo;
    f3 F0;
    float alpha;
    float cc;
    float ccAlpha;
};
inline UpsLobes ups_init(const sb_material& m)
{
    UpsLobes L;
    const f3 base{ m.base_color[0], m.base_color[1], m.base_color[2] };
    const float metallic = saturate(m.metallic);
    const float rough = saturate(m.roughness);
    L.alpha = std::fmax(rough * rough, 1e-3f);
    const float r0 = (1.0f - m.ior) / (1.0f + m.ior);
    const float f0d = r0 * r0;
    if (m.use_specular_workflow)
    {
        L.F0 = f3{ m.specular_color[0], m.specular_color[1], m.specular_color[2] };
        L.diffAlbedo = base;
    }
    else
    {
        L.F0 = lerp(mk3(f0d), base, metallic);
        L.diffAlbedo = base * (1.0f - metallic);
    }
    L.cc = saturate(m.clearcoat);
    const float ccr = saturate(m.clearcoat_roughness);
    L.ccAlpha = std::fmax(ccr * ccr, 1e-3f);
    return L;
}
struct UpsWeights
{
    float pc, ps, pd; // lobe selection probabilities (sum 1), all 0 -> absorb
    float att, wd;
};
inline UpsWeights ups_weights(const UpsLobes& L, float nk1)
{
    UpsWeights w;
    const float fc = L.cc * schlick(0.04f, nk1);
    w.att = 1.0f - fc;
    const float f0s = (L.F0.x + L.F0.y + L.F0.z) * (1.0f / 3.0f);
    w.wd = 1.0f - schlick(f0s, nk1);
    const float ws = w.att * luminance(schlick3(L.F0, nk1));
    const float wdl = w.att * w.wd * luminance(L.diffAlbedo);
    const float sum = fc + ws + wdl;
    if (!(sum > 0.0f))
    {
        w.pc = w.ps = w.pd = 0.0f;
        return w;
    }
    const float inv = 1.0f / sum;
    w.pc = fc * inv;
    w.ps = ws * inv;
    w.pd = wdl * inv;
    return w;
}
// evaluate for k1,k2 both above the surface (nk1 > 0, nk2 > 0)
inline BsdfEval ups_eval_core(const UpsLobes& L, const UpsWeights& w, const f3& n, const f3& k1, const f3& k2,
                              float nk1, float nk2)
{
    BsdfEval e;
    const f3 h = normalize(k1 + k2);
    const float nh = std::fmax(dot(n, h), 0.0f);
    const float hk = std::fmax(dot(k1, h), 0.0f);
    const f3 nxh = cross(n, h);
    const float sin2h = dot(nxh, nxh);
    const float ds = ggx_d(L.alpha, nh, sin2h);
    const float g1v = smith_g1(L.alpha, nk1);
    const float g1l = smith_g1(L.alpha, nk2);
    const f3 fs = schlick3(L.F0, hk);
    const float specScalar = ds * g1v * g1l / (4.0f * nk1 * nk2);
    f3 glossy = fs * (w.att * specScalar);
    float pdf = w.ps * (g1v * ds / (4.0f * nk1)) + w.pd * (nk2 / kPi);
    if (L.cc > 0.0f)
    {
        const float dc = ggx_d(L.ccAlpha, nh, sin2h);
        const float c1v = smith_g1(L.ccAlpha, nk1);
        const float c1l = smith_g1(L.ccAlpha, nk2);
        const float fc = L.cc * schlick(0.04f, hk);
        glossy += mk3(fc * dc * c1v * c1l / (4.0f * nk1 * nk2));
        pdf += w.pc * (c1v * dc / (4.0f * nk1));
    }
    e.glossy = glossy * nk2;
    e.diffuse = L.diffAlbedo * (w.att * w.wd * (nk2 / kPi));
    e.pdf = pdf;
    return e;
}

// ---- protocol entry points ----------------------------------------------------------------------

// mdlcode_evaluate stand-in.  n = shading normal, ng = geometric normal (both already flipped by
// `inside`, closest_hit.cu:405-406), k1 = -ray_dir, k2 = direction to the light.
inline BsdfEval bsdf_evaluate(const sb_material& m, const f3& n, const f3& ng, const f3& tangent, const f3& k1, const f3& k2)
{
    BsdfEval e{ mk3(0.0f), mk3(0.0f), 0.0f };
    if (m.model == SB_MATERIAL_HAIR)
    {
        // fibre scattering (oracle/hair.h, written from the papers in double precision): everything in the glossy slot
        const hair::Model H(m);
        const hair::Frame F(tangent, n);
        double wo[3], wi[3], fc[3], pdf;
        F.to_local(k1, wo);
        F.to_local(k2, wi);
        H.eval(wo, wi, fc, pdf);
        e.glossy = f3{ float(fc[0]), float(fc[1]), float(fc[2]) };
        e.pdf = float(pdf);
        return e;
    }
    const float nk1 = dot(n, k1);
    const float nk2 = dot(n, k2);
    if (!(nk1 > 0.0f) || !(nk2 > 0.0f) || !(dot(ng, k1) > 0.0f) || !(dot(ng, k2) > 0.0f))
    {
        return e;
    }
    if (m.model == SB_MATERIAL_USD_PREVIEW_SURFACE)
    {
        const UpsLobes L = ups_init(m);
        const UpsWeights w = ups_weights(L, nk1);
        if (w.pc + w.ps + w.pd <= 0.0f)
            return e;
        return ups_eval_core(L, w, n, k1, k2, nk1, nk2);
    }
    // SB_MATERIAL_DIFFUSE: Lambert
    const f3 c{ m.base_color[0], m.base_color[1], m.base_color[2] };
    e.diffuse = c * (nk2 / kPi);
    e.pdf = nk2 / kPi;
    return e;
}

// mdlcode_sample stand-in.  xi = (z1..z4) of closest_hit.cu:510-519.
inline BsdfSample bsdf_sample(const sb_material& m, const f3& n, const f3& ng, const f3& tangent, const f3& k1, const f4& xi)
{
    BsdfSample s;
    s.k2 = mk3(0.0f);
    s.bsdf_over_pdf = mk3(0.0f);
    s.pdf = 0.0f;
    s.event = EV_ABSORB;
    if (m.model == SB_MATERIAL_HAIR)
    {
        const hair::Model H(m);
        const hair::Frame F(tangent, n);
        double wo[3], wi[3], fc[3], pdf;
        F.to_local(k1, wo);
        const double u[4] = { xi.x, xi.y, xi.z, xi.w };
        H.sample(wo, u, wi);
        H.eval(wo, wi, fc, pdf);
        if (!(pdf > 0.0))
            return s;
        s.k2 = F.to_world(wi);
        s.pdf = float(pdf);
        s.bsdf_over_pdf = f3{ float(fc[0] / pdf), float(fc[1] / pdf), float(fc[2] / pdf) };
        // transmission = leaving through the other side of the geometric surface (inside toggles, closest_hit.cu:591-600)
        s.event = EV_GLOSSY | (dot(ng, s.k2) >= 0.0f ? EV_REFLECTION : EV_TRANSMISSION);
        return s;
    }
    const float nk1 = dot(n, k1);
    if (!(nk1 > 0.0f) || !(dot(ng, k1) > 0.0f))
    {
        return s; // seen from below: absorb (no facing test in the reference, quirk Q13)
    }
    if (m.model == SB_MATERIAL_USD_PREVIEW_SURFACE)
    {
        const UpsLobes L = ups_init(m);
        const UpsWeights w = ups_weights(L, nk1);
        if (w.pc + w.ps + w.pd <= 0.0f)
            return s;
        int event;
        f3 k2;
        if (xi.z < w.pc + w.ps)
        {
            const float a = (xi.z < w.pc) ? L.ccAlpha : L.alpha;
            const f3 h = ggx_sample_vndf(a, n, k1, xi.x, xi.y);
            k2 = 2.0f * dot(k1, h) * h - k1;
            event = EV_GLOSSY | EV_REFLECTION;
        }
        else
        {
            k2 = cosine_hemisphere(xi.x, xi.y, n);
            event = EV_DIFFUSE | EV_REFLECTION;
        }
        const float nk2 = dot(n, k2);
        if (!(nk2 > 0.0f) || !(dot(ng, k2) > 0.0f))
            return s;
        const BsdfEval e = ups_eval_core(L, w, n, k1, k2, nk1, nk2);
        if (!(e.pdf > 0.0f))
            return s;
        s.k2 = k2;
        s.pdf = e.pdf;
        s.bsdf_over_pdf = (e.diffuse + e.glossy) / e.pdf;
        s.event = event;
        return s;
    }
    const f3 k2 = cosine_hemisphere(xi.x, xi.y, n);
    const float nk2 = dot(n, k2);
    if (!(nk2 > 0.0f) || !(dot(ng, k2) > 0.0f))
        return s;
    s.k2 = k2;
    s.pdf = nk2 / kPi;
    s.bsdf_over_pdf = f3{ m.base_color[0], m.base_color[1], m.base_color[2] };
    s.event = EV_DIFFUSE | EV_REFLECTION;
    return s;
}

} // namespace orc
