// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// Pins the oracle against the REFERENCE'S OWN CODE: this file #includes the reference headers that
// compile on the host, straight from the read-only reference tree (never copied into this repo):
//     src/render/optix/RandomSampler.h, include/render/Lights.h,
//     src/render/optix/postprocessing/Utils.h, cuda/curve.h, sutil/*
// and prints golden input/output vectors as JSON (floats as IEEE-754 bit patterns).  The output is
// committed as tests/golden/ref_vectors.json together with this generator; tests compare the oracle
// (CPU, `-m "not gpu"`) and the CUDA kernels (`-m gpu`) against it, so the pin travels to the GPU box
// where /root/reference does not exist.  Build + run: `make -C oracle golden` (needs $STRELKA_REF_DIR
// or /root/reference).  Output binary goes to oracle/_ref/ only.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <vector>
#include <string>

#include <cuda_runtime.h> // vector types + make_float3 (host side)

using std::isnan;
using std::max;
using std::min;
// Under nvcc the global namespace holds float overloads of the C math functions
// (crt/math_functions.hpp), so `acos(float)`, `cos(float)`, `sqrt(float)`, `fabs(float)` in the
// reference's headers are FLOAT calls in the device code that actually runs.  A plain g++ host build
// would silently pick the double versions; pull in the float overloads to keep device semantics.
using std::acos;
using std::cos;
using std::sin;
using std::sqrt;
using std::fabs;

#include <sutil/vec_math.h>
#include <sutil/vec_math_adv.h>
#include "RandomSampler.h"
#include "Lights.h"
#include <postprocessing/Utils.h>
#include <cuda/curve.h>

static uint32_t bits(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}

struct Rng
{
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint32_t u32()
    {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        return uint32_t(s >> 16);
    }
    float f01() { return (u32() >> 8) * (1.0f / 16777216.0f); }
    float range(float a, float b) { return a + (b - a) * f01(); }
};

static void print_u32_array(const char* name, const std::vector<uint32_t>& v, bool last = false)
{
    std::printf("  \"%s\": [", name);
    for (size_t i = 0; i < v.size(); ++i)
        std::printf("%s%u", i ? "," : "", v[i]);
    std::printf("]%s\n", last ? "" : ",");
}

template <SampleDimension D>
static float rnd_dim(SamplerState& s)
{
    return random<D>(s);
}
static float rnd_any(uint32_t dim, SamplerState& s)
{
    switch (dim)
    {
    case 0: return rnd_dim<SampleDimension::ePixelX>(s);
    case 1: return rnd_dim<SampleDimension::ePixelY>(s);
    case 2: return rnd_dim<SampleDimension::eLightId>(s);
    case 3: return rnd_dim<SampleDimension::eLightPointX>(s);
    case 4: return rnd_dim<SampleDimension::eLightPointY>(s);
    case 5: return rnd_dim<SampleDimension::eBSDF0>(s);
    case 6: return rnd_dim<SampleDimension::eBSDF1>(s);
    case 7: return rnd_dim<SampleDimension::eBSDF2>(s);
    case 8: return rnd_dim<SampleDimension::eBSDF3>(s);
    default: return rnd_dim<SampleDimension::eRussianRoulette>(s);
    }
}

static UniformLight make_rect(Rng& r)
{
    // a rectangle with orthogonal edges, random pose: points (+,+),(-,+),(-,-),(+,-) like scene.cpp:363-366
    float3 c = make_float3(r.range(-2, 2), r.range(-2, 2), r.range(-2, 2));
    float3 a = normalize(make_float3(r.range(-1, 1), r.range(-1, 1), r.range(-1, 1)));
    float3 t = make_float3(r.range(-1, 1), r.range(-1, 1), r.range(-1, 1));
    float3 b = normalize(cross(a, t));
    const float w = r.range(0.1f, 2.0f), h = r.range(0.1f, 2.0f);
    UniformLight l = {};
    float3 p0 = c + a * (0.5f * w) + b * (0.5f * h);
    float3 p1 = c - a * (0.5f * w) + b * (0.5f * h);
    float3 p2 = c - a * (0.5f * w) - b * (0.5f * h);
    float3 p3 = c + a * (0.5f * w) - b * (0.5f * h);
    l.points[0] = make_float4(p0.x, p0.y, p0.z, 1);
    l.points[1] = make_float4(p1.x, p1.y, p1.z, 1);
    l.points[2] = make_float4(p2.x, p2.y, p2.z, 1);
    l.points[3] = make_float4(p3.x, p3.y, p3.z, 1);
    l.color = make_float4(r.range(0.5f, 50), r.range(0.5f, 50), r.range(0.5f, 50), 1);
    l.type = 0;
    return l;
}

static void push_light(std::vector<uint32_t>& v, const UniformLight& l)
{
    const float* f = reinterpret_cast<const float*>(&l);
    for (int i = 0; i < 24; ++i)
        v.push_back(bits(f[i]));
    v.push_back(uint32_t(l.type));
    v.push_back(bits(l.halfAngle));
    v.push_back(0);
    v.push_back(0);
}
static void push_sample(std::vector<uint32_t>& v, const LightSampleData& s)
{
    const float o[12] = { s.pointOnLight.x, s.pointOnLight.y, s.pointOnLight.z, s.pdf, s.normal.x, s.normal.y,
                          s.normal.z, s.area, s.L.x, s.L.y, s.L.z, s.distToLight };
    for (float f : o)
        v.push_back(bits(f));
}

int main()
{
    std::printf("{\n");
    std::printf("  \"generator\": \"oracle/ref_crosscheck.cpp against the reference headers RandomSampler.h, Lights.h, "
                "postprocessing/Utils.h, cuda/curve.h\",\n");
    // ---- sampler -------------------------------------------------------------------------------
    {
        std::vector<uint32_t> ints = { EncodeMorton2(3, 5), EncodeMorton2(1023, 767), hash(52), hash_combine(hash(52), 3),
                                       laine_karras_permutation(1, 2), nested_uniform_scramble(12345, hash(52)),
                                       sobol_uint(5, 2) };
        print_u32_array("sampler_ints", ints);
        std::vector<uint32_t> table;
        for (int d = 0; d < 5; ++d)
            for (int b = 0; b < 32; ++b)
                table.push_back(sb_matrix[d][b]);
        print_u32_array("sobol_matrix", table);
        Rng r(1);
        std::vector<uint32_t> in, out;
        const uint32_t fixed[][4] = { { 0, 0, 0, 256 },          { 1, 0, 0, 256 },          { 3, 5, 7, 256 },
                                      { 1023, 767, 2047, 2048 }, { 3839, 2159, 4095, 4096 }, { 1919, 1079, 2047, 2048 } };
        for (auto& f : fixed)
        {
            for (uint32_t depth = 0; depth < 2; ++depth)
            {
                for (uint32_t dim = 0; dim < 10; ++dim)
                {
                    SamplerState s = initSampler(f[0], f[1], 0, f[2], f[3], 52u);
                    s.depth = depth;
                    in.insert(in.end(), { f[0], f[1], f[2], f[3], depth, dim });
                    out.push_back(bits(rnd_any(dim, s)));
                }
            }
        }
        for (int i = 0; i < 2000; ++i)
        {
            const uint32_t maxs = 1u << (r.u32() % 13);
            const uint32_t x = r.u32() % 3840, y = r.u32() % 2160, smp = r.u32() % maxs, depth = r.u32() % 8, dim = r.u32() % 10;
            SamplerState s = initSampler(x, y, 0, smp, maxs, 52u);
            s.depth = depth;
            in.insert(in.end(), { x, y, smp, maxs, depth, dim });
            out.push_back(bits(rnd_any(dim, s)));
        }
        print_u32_array("sampler_in", in);
        print_u32_array("sampler_out", out);
    }
    // ---- lights --------------------------------------------------------------------------------
    {
        Rng r(2);
        std::vector<uint32_t> lights, hp, uv, sUniform, sSphQuad, sSphere, sDistant, pdfs;
        for (int i = 0; i < 600; ++i)
        {
            UniformLight l = make_rect(r);
            const int kind = i % 3;
            if (kind == 1)
            {
                l.type = 2;
                l.points[0] = make_float4(r.range(0.05f, 1.0f), 0, 0, 0);
                l.points[1] = make_float4(r.range(-2, 2), r.range(-2, 2), r.range(-2, 2), 1);
            }
            else if (kind == 2)
            {
                l.type = 3;
                l.halfAngle = r.range(0.001f, 0.5f);
                float3 n = normalize(make_float3(r.range(-1, 1), r.range(-1, 1), r.range(-1, 1)));
                l.normal = make_float4(n.x, n.y, n.z, 0);
            }
            const float3 p = make_float3(r.range(-3, 3), r.range(-3, 3), r.range(-3, 3));
            const float2 u = make_float2(r.f01(), r.f01());
            push_light(lights, l);
            hp.insert(hp.end(), { bits(p.x), bits(p.y), bits(p.z) });
            uv.insert(uv.end(), { bits(u.x), bits(u.y) });
            if (l.type == 0)
            {
                push_sample(sUniform, SampleRectLightUniform(l, u, p));
                push_sample(sSphQuad, SampleRectLight(l, u, p));
            }
            else if (l.type == 2)
            {
                push_sample(sSphere, SampleSphereLight(l, u, p));
            }
            else
            {
                push_sample(sDistant, SampleDistantLight(l, u, p));
            }
            // emitter-hit side: a point on/near the light as "lightHitPoint", p as ray origin
            float3 lh;
            if (l.type == 0)
                lh = make_float3(l.points[0]) + (make_float3(l.points[1]) - make_float3(l.points[0])) * u.x +
                     (make_float3(l.points[3]) - make_float3(l.points[0])) * u.y;
            else
                lh = make_float3(l.points[1]) + make_float3(0.3f, -0.2f, 0.1f);
            const float pdf = getLightPdf(l, lh, p);
            const float3 n = calcLightNormal(l, lh);
            pdfs.insert(pdfs.end(), { bits(lh.x), bits(lh.y), bits(lh.z), bits(pdf), bits(n.x), bits(n.y), bits(n.z) });
        }
        print_u32_array("light_structs", lights);
        print_u32_array("light_hit_points", hp);
        print_u32_array("light_u", uv);
        print_u32_array("light_sample_rect_uniform", sUniform);
        print_u32_array("light_sample_rect_sphquad", sSphQuad);
        print_u32_array("light_sample_sphere", sSphere);
        print_u32_array("light_sample_distant", sDistant);
        print_u32_array("light_pdf_normal", pdfs);
        std::vector<uint32_t> mis;
        for (int i = 0; i < 64; ++i)
        {
            const float a = r.range(0.001f, 50), b = r.range(0.0f, 50);
            mis.insert(mis.end(), { bits(a), bits(b), bits(misWeightBalance(a, b)) });
        }
        print_u32_array("mis_balance", mis);
    }
    // ---- tonemap / accumulate ------------------------------------------------------------------------
    {
        Rng r(3);
        std::vector<uint32_t> v;
        for (int i = 0; i < 256; ++i)
        {
            const float3 c = make_float3(r.range(0, 100), r.range(0, 100), r.range(0, 100));
            const float3 c2 = make_float3(r.range(0, 100), r.range(0, 100), r.range(0, 100));
            const float ev = (i % 2) ? 6.25e-4f : r.range(1e-4f, 1.0f);
            const float3 e = make_float3(ev);
            const uint32_t sub = 1 + r.u32() % 4096;
            const float3 t = tonemap(c, e);
            const float3 it = inverseTonemap(t, e);
            // accumulate(), OptixRender.cu:60-78, restated with the reference's own helpers
            const float a = 1.0f / static_cast<float>(sub + 1);
            const float3 acc = inverseTonemap(lerp(tonemap(c, e), tonemap(c2, e), a), e);
            v.insert(v.end(), { bits(c.x), bits(c.y), bits(c.z), bits(c2.x), bits(c2.y), bits(c2.z), bits(ev), sub, bits(t.x),
                                bits(t.y), bits(t.z), bits(it.x), bits(it.y), bits(it.z), bits(acc.x), bits(acc.y), bits(acc.z) });
        }
        print_u32_array("tonemap", v);
    }
    // ---- curves ----------------------------------------------------------------------------------
    {
        Rng r(4);
        std::vector<uint32_t> v;
        for (int i = 0; i < 256; ++i)
        {
            float4 q[4];
            float3 p = make_float3(r.range(-1, 1), r.range(-1, 1), r.range(-1, 1));
            for (int k = 0; k < 4; ++k)
            {
                p += make_float3(r.range(0.1f, 0.5f), r.range(-0.3f, 0.3f), r.range(-0.3f, 0.3f));
                q[k] = make_float4(p.x, p.y, p.z, r.range(0.005f, 0.1f));
            }
            CubicInterpolator ci;
            ci.initializeFromBSpline(q);
            const float u = (i < 8) ? ((i % 2) ? 1.0f : 0.0f) : r.f01();
            const float4 pos = ci.position4(u);
            const float4 vel = ci.velocity4(u);
            const float3 tg = curveTangent(ci, u);
            // a point near the surface at parameter u
            float3 side = normalize(cross(make_float3(vel), make_float3(0.3f, 0.5f, -0.8f)));
            float3 ps = make_float3(pos) + side * (pos.w * r.range(0.9f, 1.1f)) + make_float3(vel) * r.range(-0.01f, 0.01f);
            const float3 ps0 = ps;
            const float3 n = surfaceNormal(ci, u, ps);
            for (int k = 0; k < 4; ++k)
                v.insert(v.end(), { bits(q[k].x), bits(q[k].y), bits(q[k].z), bits(q[k].w) });
            v.insert(v.end(), { bits(u), bits(ps0.x), bits(ps0.y), bits(ps0.z) });
            v.insert(v.end(), { bits(pos.x), bits(pos.y), bits(pos.z), bits(pos.w), bits(vel.x), bits(vel.y), bits(vel.z), bits(vel.w),
                                bits(tg.x), bits(tg.y), bits(tg.z), bits(n.x), bits(n.y), bits(n.z), bits(ps.x), bits(ps.y), bits(ps.z) });
        }
        print_u32_array("curve", v, true);
    }
    std::printf("}\n");
    return 0;
}
