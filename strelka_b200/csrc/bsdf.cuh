// Device BSDFs (closed forms) behind the reference's MDL call protocol (closest_hit.cu:474-605):
// sample(k1, xi[4]) -> k2, bsdf_over_pdf, pdf, event bits; evaluate(k1, k2) -> bsdf_diffuse,
// bsdf_glossy (cosine INCLUDED) and pdf.  PARITY UNPINNED against the reference (its BSDF arithmetic
// lives in NVIDIA's closed MDL SDK); the definitions are this repository's
// (DESIGN.md "Materials") and are checked against the CPU oracle, which states the same formulas.
#pragma once
#include "hd.cuh"
#include "hair.cuh"
#include "../../include/sb/sb_api.h"

namespace sb
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
    float3 k2;
    float3 bsdf_over_pdf;
    float pdf;
    int event;
};
struct BsdfEval
{
    float3 diffuse; // cos included
    float3 glossy; // cos included
    float pdf;
};

// Branch-light orthonormal basis around a unit vector (Duff et al. 2017)
SB_HD void onb(const float3& n, float3& t, float3& b)
{
    const float sign = copysignf(1.0f, n.z);
    const float a = -1.0f / (sign + n.z);
    const float c = n.x * n.y * a;
    t = mk3(1.0f + sign * n.x * n.x * a, sign * c, -sign * n.x);
    b = mk3(c, sign + n.y * n.y * a, -n.y);
}

SB_HD float3 cosine_hemisphere(float u1, float u2, const float3& n)
{
    const float r = sqrtf(u1);
    float sphi, cphi;
    sincos2pi(u2, sphi, cphi);
    const float x = r * cphi;
    const float y = r * sphi;
    const float z = sqrtf(fmaxf(0.0f, 1.0f - u1));
    float3 t, b;
    onb(n, t, b);
    return normalize(x * t + y * b + z * n);
}

SB_HD float luminance(const float3& c)
{
    return dot(c, mk3(0.299f, 0.587f, 0.114f));
}
SB_HD float pow5(float x)
{
    const float x2 = x * x;
    return x2 * x2 * x;
}
SB_HD float schlick(float f0, float c)
{
    return f0 + (1.0f - f0) * pow5(1.0f - c);
}
SB_HD float3 schlick3(const float3& f0, float c)
{
    const float w = pow5(1.0f - c);
    return f0 + (mk3(1.0f) - f0) * w;
}
// GGX normal distribution, D = a^2 / (pi (a^2 cos^2 + sin^2)^2) with sin^2 = |n x h|^2 taken from the cross product:
// the textbook form cos^2 (a^2 - 1) + 1 cancels catastrophically at the peak of a smooth lobe (a^2 = 1e-4 leaves
// three digits in fp32 -- found by the independent value pins of tests/test_bsdf_pins.py)
SB_HD float ggx_d(float a, float nh, float sin2h)
{
    const float a2 = a * a;
    const float d = a2 * nh * nh + sin2h;
    return a2 / (kPi * d * d);
}
SB_HD float smith_g1(float a, float nx)
{
    const float a2 = a * a;
    return 2.0f * nx / (nx + sqrtf(a2 + (1.0f - a2) * nx * nx));
}

// Heitz 2018 visible-normal sampling; v and result in world space around n
SB_HD float3 ggx_sample_vndf(float a, const float3& n, const float3& v, float u1, float u2)
{
    float3 t, b;
    onb(n, t, b);
    const float3 vl = mk3(dot(v, t), dot(v, b), dot(v, n));
    const float3 vh = normalize(mk3(a * vl.x, a * vl.y, vl.z));
    const float lensq = vh.x * vh.x + vh.y * vh.y;
    const float3 T1 = lensq > 0.0f ? mk3(-vh.y, vh.x, 0.0f) * (1.0f / sqrtf(lensq)) : mk3(1.0f, 0.0f, 0.0f);
    const float3 T2 = cross(vh, T1);
    const float r = sqrtf(u1);
    float sphi, cphi;
    sincos2pi(u2, sphi, cphi);
    const float t1 = r * cphi;
    float t2 = r * sphi;
    const float s = 0.5f * (1.0f + vh.z);
    t2 = (1.0f - s) * sqrtf(fmaxf(0.0f, 1.0f - t1 * t1)) + s * t2;
    const float3 nh = t1 * T1 + t2 * T2 + sqrtf(fmaxf(0.0f, 1.0f - t1 * t1 - t2 * t2)) * vh;
    const float3 hl = normalize(mk3(a * nh.x, a * nh.y, fmaxf(0.0f, nh.z)));
    return normalize(hl.x * t + hl.y * b + hl.z * n);
}

// ---- UsdPreviewSurface lobes ------------------------------------------------------------------
struct UpsLobes
{
    float3 diffAlbedo;
    float3 F0;
    float alpha;
    float cc;
    float ccAlpha;
};
SB_HD UpsLobes ups_init(const sb_material& m, const float3& base)
{
    UpsLobes L;
    const float metallic = saturate(m.metallic);
    const float rough = saturate(m.roughness);
    L.alpha = fmaxf(rough * rough, 1e-3f);
    const float r0 = (1.0f - m.ior) / (1.0f + m.ior);
    const float f0d = r0 * r0;
    if (m.use_specular_workflow)
    {
        L.F0 = mk3(m.specular_color[0], m.specular_color[1], m.specular_color[2]);
        L.diffAlbedo = base;
    }
    else
    {
        L.F0 = lerp(mk3(f0d), base, metallic);
        L.diffAlbedo = base * (1.0f - metallic);
    }
    L.cc = saturate(m.clearcoat);
    const float ccr = saturate(m.clearcoat_roughness);
    L.ccAlpha = fmaxf(ccr * ccr, 1e-3f);
    return L;
}
struct UpsWeights
{
    float pc, ps, pd; // lobe selection probabilities (sum 1), all 0 -> absorb
    float att, wd;
};
SB_HD UpsWeights ups_weights(const UpsLobes& L, float nk1)
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
SB_HD BsdfEval ups_eval_core(const UpsLobes& L, const UpsWeights& w, const float3& n, const float3& k1, const float3& k2,
                              float nk1, float nk2)
{
    BsdfEval e;
    const float3 h = normalize(k1 + k2);
    const float nh = fmaxf(dot(n, h), 0.0f);
    const float hk = fmaxf(dot(k1, h), 0.0f);
    const float3 nxh = cross(n, h);
    const float sin2h = dot(nxh, nxh);
    const float ds = ggx_d(L.alpha, nh, sin2h);
    const float g1v = smith_g1(L.alpha, nk1);
    const float g1l = smith_g1(L.alpha, nk2);
    const float3 fs = schlick3(L.F0, hk);
    const float specScalar = ds * g1v * g1l / (4.0f * nk1 * nk2);
    float3 glossy = fs * (w.att * specScalar);
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
// base = the material's base colour after texturing (m.base_color, or the diffuse texture's texel).
// PREVIEW = false compiles the UsdPreviewSurface model out (scenes whose materials are all diffuse), HAIR = false the
// hair fibre model (scenes without an SB_MATERIAL_HAIR material).  tangent = state.tangent_u (closest_hit.cu:485).
template <bool PREVIEW = true, bool HAIR = true>
SB_HD BsdfEval bsdf_evaluate(const sb_material& m, const float3& base, const float3& n, const float3& ng, const float3& tangent, const float3& k1,
                             const float3& k2)
{
    BsdfEval e;
    e.diffuse = mk3(0.0f);
    e.glossy = mk3(0.0f);
    e.pdf = 0.0f;
    if (HAIR && m.model == SB_MATERIAL_HAIR)
    {
        // a fibre scatters into the whole sphere: no hemisphere tests (hair.cuh)
        hair_evaluate(m, n, tangent, k1, k2, e.glossy, e.pdf);
        return e;
    }
    const float nk1 = dot(n, k1);
    const float nk2 = dot(n, k2);
    if (!(nk1 > 0.0f) || !(nk2 > 0.0f) || !(dot(ng, k1) > 0.0f) || !(dot(ng, k2) > 0.0f))
    {
        return e;
    }
    if (PREVIEW && m.model == SB_MATERIAL_USD_PREVIEW_SURFACE)
    {
        const UpsLobes L = ups_init(m, base);
        const UpsWeights w = ups_weights(L, nk1);
        if (w.pc + w.ps + w.pd <= 0.0f)
            return e;
        return ups_eval_core(L, w, n, k1, k2, nk1, nk2);
    }
    // SB_MATERIAL_DIFFUSE: Lambert
    e.diffuse = base * (nk2 / kPi);
    e.pdf = nk2 / kPi;
    return e;
}

// mdlcode_sample stand-in.  xi = (z1..z4) of closest_hit.cu:510-519.
template <bool PREVIEW = true, bool HAIR = true>
SB_HD BsdfSample bsdf_sample(const sb_material& m, const float3& base, const float3& n, const float3& ng, const float3& tangent, const float3& k1,
                             const float4& xi)
{
    BsdfSample s;
    s.k2 = mk3(0.0f);
    s.bsdf_over_pdf = mk3(0.0f);
    s.pdf = 0.0f;
    s.event = EV_ABSORB;
    if (HAIR && m.model == SB_MATERIAL_HAIR)
    {
        if (hair_sample(m, n, tangent, k1, xi, s.k2, s.bsdf_over_pdf, s.pdf))
            s.event = EV_GLOSSY | (dot(ng, s.k2) >= 0.0f ? EV_REFLECTION : EV_TRANSMISSION);
        return s;
    }
    const float nk1 = dot(n, k1);
    if (!(nk1 > 0.0f) || !(dot(ng, k1) > 0.0f))
    {
        return s; // seen from below: absorb (no facing test in the reference, quirk Q13)
    }
    if (PREVIEW && m.model == SB_MATERIAL_USD_PREVIEW_SURFACE)
    {
        const UpsLobes L = ups_init(m, base);
        const UpsWeights w = ups_weights(L, nk1);
        if (w.pc + w.ps + w.pd <= 0.0f)
            return s;
        int event;
        float3 k2;
        if (xi.z < w.pc + w.ps)
        {
            const float a = (xi.z < w.pc) ? L.ccAlpha : L.alpha;
            const float3 h = ggx_sample_vndf(a, n, k1, xi.x, xi.y);
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
    const float3 k2 = cosine_hemisphere(xi.x, xi.y, n);
    const float nk2 = dot(n, k2);
    if (!(nk2 > 0.0f) || !(dot(ng, k2) > 0.0f))
        return s;
    s.k2 = k2;
    s.pdf = nk2 / kPi;
    s.bsdf_over_pdf = base;
    s.event = EV_DIFFUSE | EV_REFLECTION;
    return s;
}

} // namespace sb
