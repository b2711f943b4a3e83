// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// A scalar, multithreaded CPU restatement of Strelka's path-tracing hot path
// (oka::Render::render() -> __raygen__rg -> optixTrace -> __closesthit__radiance / __closesthit__light
// / __miss__ms -> accumulate), written from the reference's algorithm with file:line citations.
// It is the parity checker for the CUDA backend and the timed CPU baseline of bench.py.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it;
// the product library never does.
//
// Parity status:
//   * sampler, light sampling, accumulate/tonemap, curve normals: PINNED bit-exactly / to 1 ulp
//     against the reference's own headers (oracle/ref_crosscheck.cpp, tests/golden/*.json).
//   * camera, integrator control flow, NEE / MIS / emitter-hit arithmetic, offset_ray, vertex
//     unpacking: restated from source (cited below); the reference ships no test or golden image
//     for them -> "parity unpinned" beyond the header-level pins.
//   * traversal / triangle / curve intersection (OptiX driver) and BSDF values (MDL SDK): closed
//     third-party code absent from the reference tree -> "parity unpinned" (see DESIGN.md).
#include "../include/sb/sb_api.h"
#include "vec.h"
#include "sampler.h"
#include "lights.h"
#include "curve.h"
#include "bsdf.h"
#include "texture.h"
#include "bvh2.h"

#include <atomic>
#include <thread>
#include <vector>
#include <cstdio>

using namespace orc;

namespace
{

// GEOMETRY_MASK_*, OptixRenderParams.h:9-17
constexpr uint32_t kMaskTriangle = 1, kMaskCurve = 2, kMaskLight = 4;
constexpr uint32_t kRayMaskPrimary = 255, kRayMaskShadow = 3;

struct InstanceData
{
    Affine o2w; // OptixInstance.transform / HitGroupData.object_to_world (OptixRender.cpp:438,792)
    Affine w2o; // HitGroupData.world_to_object = glm::inverse(transform) (OptixRender.cpp:794-796)
    uint32_t type, geom, material, light, mask;
};

struct WorldTri
{
    f3 v0, e1, e2;
    uint32_t prim, inst;
};
struct WorldSeg
{
    CurveSpan span; // power-basis coefficients of span k of K of the segment, world space (w = radius)
    uint32_t prim; // index into the curve's segment list (optixGetPrimitiveIndex)
    uint32_t inst;
    uint32_t firstPoint; // global index of the first control point (segmentIndices[prim])
    uint32_t k, K;
};

struct Hit
{
    float t, u, v;
    uint32_t prim, inst, kind; // kind 0 miss, 1 triangle, 2 curve
};

// double-precision inverse of an affine 3x4, rounded to float once (the CUDA host code uses the
// same formula so both sides hold identical world_to_object matrices)
Affine invert_affine(const Affine& a)
{
    const double m00 = a.m[0], m01 = a.m[1], m02 = a.m[2], tx = a.m[3];
    const double m10 = a.m[4], m11 = a.m[5], m12 = a.m[6], ty = a.m[7];
    const double m20 = a.m[8], m21 = a.m[9], m22 = a.m[10], tz = a.m[11];
    const double c00 = m11 * m22 - m12 * m21, c01 = m12 * m20 - m10 * m22, c02 = m10 * m21 - m11 * m20;
    const double det = m00 * c00 + m01 * c01 + m02 * c02;
    const double id = 1.0 / det;
    double r[9];
    r[0] = c00 * id;
    r[1] = (m02 * m21 - m01 * m22) * id;
    r[2] = (m01 * m12 - m02 * m11) * id;
    r[3] = c01 * id;
    r[4] = (m00 * m22 - m02 * m20) * id;
    r[5] = (m02 * m10 - m00 * m12) * id;
    r[6] = c02 * id;
    r[7] = (m01 * m20 - m00 * m21) * id;
    r[8] = (m00 * m11 - m01 * m10) * id;
    Affine o;
    o.m[0] = float(r[0]);
    o.m[1] = float(r[1]);
    o.m[2] = float(r[2]);
    o.m[3] = float(-(r[0] * tx + r[1] * ty + r[2] * tz));
    o.m[4] = float(r[3]);
    o.m[5] = float(r[4]);
    o.m[6] = float(r[5]);
    o.m[7] = float(-(r[3] * tx + r[4] * ty + r[5] * tz));
    o.m[8] = float(r[6]);
    o.m[9] = float(r[7]);
    o.m[10] = float(r[8]);
    o.m[11] = float(-(r[6] * tx + r[7] * ty + r[8] * tz));
    return o;
}

double affine_scale(const Affine& a)
{
    const double m00 = a.m[0], m01 = a.m[1], m02 = a.m[2];
    const double m10 = a.m[4], m11 = a.m[5], m12 = a.m[6];
    const double m20 = a.m[8], m21 = a.m[9], m22 = a.m[10];
    const double det = m00 * (m11 * m22 - m12 * m21) + m01 * (m12 * m20 - m10 * m22) + m02 * (m10 * m21 - m11 * m20);
    return std::cbrt(std::fabs(det));
}

//  unpackNormal / unpackUV, closest_hit.cu:236-254
f3 unpack_normal(uint32_t val)
{
    f3 n;
    n.z = ((val & 0xfff00000u) >> 20) / 511.99999f * 2.0f - 1.0f;
    n.y = ((val & 0x000ffc00u) >> 10) / 511.99999f * 2.0f - 1.0f;
    n.x = (val & 0x000003ffu) / 511.99999f * 2.0f - 1.0f;
    return n;
}

// offset_ray, closest_hit.cu:218-233 (Ray Tracing Gems ch. 6)
f3 offset_ray(const f3& p, const f3& n)
{
    const float origin = 1.0f / 32.0f;
    const float float_scale = 1.0f / 65536.0f;
    const float int_scale = 256.0f;
    const int32_t ix = int32_t(int_scale * n.x), iy = int32_t(int_scale * n.y), iz = int32_t(int_scale * n.z);
    const f3 pi{ i2f(f2i(p.x) + ((p.x < 0) ? -ix : ix)), i2f(f2i(p.y) + ((p.y < 0) ? -iy : iy)),
                 i2f(f2i(p.z) + ((p.z < 0) ? -iz : iz)) };
    return f3{ std::fabs(p.x) < origin ? p.x + float_scale * n.x : pi.x,
               std::fabs(p.y) < origin ? p.y + float_scale * n.y : pi.y,
               std::fabs(p.z) < origin ? p.z + float_scale * n.z : pi.z };
}

f3 interp3(const f3& a, const f3& b, const f3& c, float bx, float by)
{
    // interpolateAttrib, closest_hit.cu:199-205
    return a * (1.0f - bx - by) + b * bx + c * by;
}

// tonemap / inverseTonemap, postprocessing/Utils.h:5-15
f3 tonemap3(f3 c, const f3& e)
{
    c = c * e;
    return c / (c + mk3(1.0f));
}
f3 inverse_tonemap3(const f3& c, const f3& e)
{
    return c / (e - c * e);
}

} // namespace

struct orc_scene
{
    std::vector<sb_vertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<sb_mesh> meshes;
    std::vector<sb_curve> curves;
    std::vector<f3> curvePoints;
    std::vector<float> curveRadii;
    std::vector<uint32_t> curveVertexCounts;
    std::vector<sb_light> lights;
    std::vector<sb_material> materials;
    std::vector<Texture> textures;
    std::vector<InstanceData> instances;
    std::vector<WorldTri> tris;
    std::vector<WorldSeg> segs;
    Bvh2 triBvh, segBvh;
    uint32_t curveSplit = 8;
};

namespace
{

void build_scene(orc_scene& S, const sb_scene_view& v)
{
    S.vertices.assign(v.vertices, v.vertices + v.num_vertices);
    S.indices.assign(v.indices, v.indices + v.num_indices);
    S.meshes.assign(v.meshes, v.meshes + v.num_meshes);
    S.textures.clear();
    for (uint32_t i = 0; v.textures && i < v.num_textures; ++i)
    {
        Texture t;
        t.width = v.textures[i].width;
        t.height = v.textures[i].height;
        t.rgba.assign(v.textures[i].pixels, v.textures[i].pixels + size_t(t.width) * t.height * 4);
        S.textures.push_back(std::move(t));
    }
    S.curves.assign(v.curves, v.curves + v.num_curves);
    S.curvePoints.resize(v.num_curve_points);
    for (uint64_t i = 0; i < v.num_curve_points; ++i)
        S.curvePoints[i] = f3{ v.curve_points[3 * i], v.curve_points[3 * i + 1], v.curve_points[3 * i + 2] };
    S.curveRadii.assign(v.curve_widths, v.curve_widths + v.num_curve_widths);
    S.curveVertexCounts.assign(v.curve_vertex_counts, v.curve_vertex_counts + v.num_curve_vertex_counts);
    S.lights.assign(v.lights, v.lights + v.num_lights);
    S.materials.assign(v.materials, v.materials + v.num_materials);
    if (S.materials.empty())
    {
        sb_material m{};
        m.model = SB_MATERIAL_DIFFUSE;
        m.base_color[0] = m.base_color[1] = m.base_color[2] = 1.0f;
        S.materials.push_back(m);
    }
    S.instances.resize(v.num_instances);
    std::vector<Aabb> triBoxes, segBoxes;
    for (uint32_t i = 0; i < v.num_instances; ++i)
    {
        const sb_instance& in = v.instances[i];
        InstanceData& d = S.instances[i];
        // glm column-major -> row-major 3x4 (glm::rowMajor4, OptixRender.cpp:438)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c)
                d.o2w.m[r * 4 + c] = in.transform[c * 4 + r];
        d.w2o = invert_affine(d.o2w);
        d.type = in.type;
        d.geom = in.geom_id;
        // OptixRender.cpp:766: unknown material -> default (slot 0)
        d.material = (in.material_id == 0xffffffffu || in.material_id >= S.materials.size()) ? 0u : in.material_id;
        d.light = in.light_id;
        d.mask = in.type == SB_INSTANCE_MESH ? kMaskTriangle : (in.type == SB_INSTANCE_CURVE ? kMaskCurve : kMaskLight);
        if (in.type == SB_INSTANCE_MESH || in.type == SB_INSTANCE_LIGHT)
        {
            if (in.geom_id >= S.meshes.size())
                continue;
            const sb_mesh& m = S.meshes[in.geom_id];
            const uint32_t ntri = m.count / 3;
            for (uint32_t t = 0; t < ntri; ++t)
            {
                f3 p[3];
                for (int k = 0; k < 3; ++k)
                {
                    const sb_vertex& vx = S.vertices[m.vb_offset + S.indices[m.index + 3 * t + k]];
                    p[k] = xform_point(d.o2w, f3{ vx.pos[0], vx.pos[1], vx.pos[2] });
                }
                WorldTri wt;
                wt.v0 = p[0];
                wt.e1 = p[1] - p[0];
                wt.e2 = p[2] - p[0];
                wt.prim = t;
                wt.inst = i;
                S.tris.push_back(wt);
                Aabb b;
                b.reset();
                b.grow(p[0]);
                b.grow(p[1]);
                b.grow(p[2]);
                triBoxes.push_back(b);
            }
        }
        else if (in.type == SB_INSTANCE_CURVE)
        {
            if (in.geom_id >= S.curves.size())
                continue;
            const sb_curve& c = S.curves[in.geom_id];
            const float scale = float(affine_scale(d.o2w));
            // segment list of OptiXRender::createCurve, OptixRender.cpp:232-245
            uint32_t offsetInside = 0, prim = 0;
            for (uint32_t ci = 0; ci < c.vertex_counts_count; ++ci)
            {
                const uint32_t ncp = S.curveVertexCounts[c.vertex_counts_start + ci];
                const int nseg = int(ncp) - 3;
                for (int s = 0; s < nseg; ++s)
                {
                    const uint32_t first = c.points_start + offsetInside + uint32_t(s);
                    f4 q[4];
                    for (int k = 0; k < 4; ++k)
                    {
                        const f3 pw = xform_point(d.o2w, S.curvePoints[first + k]);
                        const float r = (first + k < S.curveRadii.size() ? S.curveRadii[first + k] : 0.0f) * scale;
                        q[k] = mk4(pw, r);
                    }
                    for (uint32_t k = 0; k < S.curveSplit; ++k)
                    {
                        WorldSeg ws;
                        ws.span = curve_span(q, k, S.curveSplit);
                        ws.prim = prim;
                        ws.inst = i;
                        ws.firstPoint = first;
                        ws.k = k;
                        ws.K = S.curveSplit;
                        Aabb b;
                        curve_span_bounds(ws.span, b.lo, b.hi);
                        S.segs.push_back(ws);
                        segBoxes.push_back(b);
                    }
                    ++prim;
                }
                offsetInside += ncp;
            }
        }
    }
    S.triBvh.build(triBoxes);
    S.segBvh.build(segBoxes);
}

// Moller-Trumbore on precomputed edges, no culling (OPTIX_RAY_FLAG_NONE, OptixRender.cu:125).
// Expression tree is part of the arithmetic contract with the CUDA kernel (fma-based dot/cross).
inline bool intersect_tri(const WorldTri& T, const f3& o, const f3& d, float tmin, float tmax, float& t, float& u, float& v)
{
    const f3 p = cross_fma(d, T.e2);
    const float det = dot_fma(T.e1, p);
    if (det == 0.0f)
        return false;
    const float inv = 1.0f / det;
    const f3 tv = o - T.v0;
    u = dot_fma(tv, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f))
        return false;
    const f3 q = cross_fma(tv, T.e1);
    v = dot_fma(d, q) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f))
        return false;
    t = dot_fma(T.e2, q) * inv;
    return t > tmin && t <= tmax;
}

Hit trace_closest(const orc_scene& S, const f3& o, const f3& d, float tmin, float tmaxIn, uint32_t rayMask,
                  uint64_t* nTriTests = nullptr)
{
    Hit h{ tmaxIn, 0, 0, 0, 0, 0 };
    uint32_t bestId = 0xffffffffu;
    S.triBvh.traverse(o, d, tmin, tmaxIn, [&](uint32_t id, float& tmax) {
        const WorldTri& T = S.tris[id];
        if (!(S.instances[T.inst].mask & rayMask))
            return false;
        if (nTriTests)
            ++*nTriTests;
        float t, u, v;
        if (intersect_tri(T, o, d, tmin, tmax, t, u, v))
        {
            if (t < h.t || h.kind == 0 || (t == h.t && id < bestId))
            {
                h = Hit{ t, u, v, T.prim, T.inst, 1 };
                bestId = id;
                tmax = t;
            }
        }
        return false;
    });
    if (rayMask & kMaskCurve)
    {
        float tmaxC = h.kind ? h.t : tmaxIn;
        uint32_t bestSeg = 0xffffffffu;
        bool curveWon = false;
        Hit hc = h;
        S.segBvh.traverse(o, d, tmin, tmaxC, [&](uint32_t id, float& tmax) {
            const WorldSeg& W = S.segs[id];
            CurveHit ch{ false, 0.0f, 0.0f };
            ch.hit = intersect_round_cubic_f32(W.span.c, o, d, tmin, tmax, ch.t, ch.u);
            if (ch.hit && (ch.t < tmax || (curveWon && ch.t == tmax && id < bestSeg)))
            {
                const float useg = (float(W.k) + ch.u) * (1.0f / float(W.K)); // span-local -> segment parameter
                hc = Hit{ ch.t, useg, 0.0f, W.prim, W.inst, 2 };
                bestSeg = id;
                curveWon = true;
                tmax = ch.t;
            }
            return false;
        });
        if (curveWon)
            h = hc;
    }
    if (h.kind == 0)
        h.t = 0.0f;
    return h;
}

bool trace_any(const orc_scene& S, const f3& o, const f3& d, float tmin, float tmaxIn, uint32_t rayMask)
{
    bool occluded = false;
    S.triBvh.traverse(o, d, tmin, tmaxIn, [&](uint32_t id, float& tmax) {
        const WorldTri& T = S.tris[id];
        if (!(S.instances[T.inst].mask & rayMask))
            return false;
        float t, u, v;
        if (intersect_tri(T, o, d, tmin, tmax, t, u, v))
        {
            occluded = true;
            return true;
        }
        return false;
    });
    if (occluded || !(rayMask & kMaskCurve))
        return occluded;
    S.segBvh.traverse(o, d, tmin, tmaxIn, [&](uint32_t id, float& tmax) {
        float tc, uc;
        if (intersect_round_cubic_f32(S.segs[id].span.c, o, d, tmin, tmax, tc, uc))
        {
            occluded = true;
            return true;
        }
        return false;
    });
    return occluded;
}

struct RenderParams
{
    const orc_scene* scene;
    sb_settings st;
    float clipToView[16], viewToWorld[16];
    uint32_t width, height;
    f3 exposure;
};

// PerRayData, OptixRenderParams.h:79-93
struct Prd
{
    Sampler sampler;
    uint32_t depth;
    f3 radiance, throughput, origin, dir;
    bool inside, specularBounce;
    float lastBsdfPdf;
    uint8_t firstEventType; // EventType: 0 undef, 1 absorb, 2 diffuse, 3 specular
};

struct Counters
{
    uint64_t paths = 0, radianceRays = 0, shadowRays = 0;
};

// row-major 4x4 * float4 (sutil::Matrix4x4 operator*, sutil/Matrix.h)
inline f4 mat4_mul(const float* m, const f4& v)
{
    return f4{ m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w, m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * v.w,
               m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * v.w, m[12] * v.x + m[13] * v.y + m[14] * v.z + m[15] * v.w };
}

// generateCameraRay, OptixRender.cu:38-58
void camera_ray(const RenderParams& P, uint32_t px, uint32_t py, Sampler& s, f3& origin, f3& dir)
{
    const float jx = rnd(kPixelX, s), jy = rnd(kPixelY, s);
    const float posx = float(px) + jx, posy = float(py) + jy;
    const float ndcx = (posx / float(P.width)) * 2.0f - 1.0f;
    const float ndcy = (posy / float(P.height)) * 2.0f - 1.0f;
    const f4 clip{ ndcx, ndcy, 1.0f, 1.0f };
    const f4 vs = mat4_mul(P.clipToView, clip);
    const f4 wdir = mat4_mul(P.viewToWorld, f4{ vs.x, vs.y, vs.z, 0.0f });
    origin = mk3(mat4_mul(P.viewToWorld, f4{ 0.0f, 0.0f, 0.0f, 1.0f }));
    dir = normalize(mk3(wdir));
}

// __closesthit__light, OptixRender.cu:315-341
void hit_light(const RenderParams& P, const InstanceData& inst, const f3& rayO, const f3& rayD, float t, Prd& prd)
{
    const orc_scene& S = *P.scene;
    if (inst.light < S.lights.size())
    {
        const sb_light& l = S.lights[inst.light];
        const f3 hitPoint = rayO + t * rayD;
        const f3 ln = light_normal(l, hitPoint);
        const f3 color{ l.color[0], l.color[1], l.color[2] };
        if (-dot(rayD, ln) > 0.0f)
        {
            if (prd.depth == 0 || prd.specularBounce)
            {
                prd.radiance += prd.throughput * color * -dot(rayD, ln);
            }
            else
            {
                const float lightPdf = light_pdf(l, hitPoint, rayO) / float(uint32_t(S.lights.size()));
                const float w = mis_balance(prd.lastBsdfPdf, lightPdf);
                prd.radiance += prd.throughput * color * -dot(rayD, ln) * w;
            }
        }
    }
    prd.throughput = mk3(0.0f);
}

struct Surface
{
    f3 position, normal, geomNormal, tangent;
    f2 uv{ 0.5f, 0.5f }; // state.text_coords[0]: interpolated st for triangles, fixed 0.5 for curves (closest_hit.cu:446)
};

// unpackUV, closest_hit.cu:246-254
inline f2 unpack_uv(uint32_t val)
{
    f2 uv;
    uv.y = float((val & 0xffff0000u) >> 16) / 16383.99999f * 20.0f - 10.0f;
    uv.x = float(val & 0x0000ffffu) / 16383.99999f * 20.0f - 10.0f;
    return uv;
}

// fillTriangleGeomData, closest_hit.cu:365-421
Surface tri_surface(const orc_scene& S, const InstanceData& inst, const Hit& h, bool inside)
{
    const sb_mesh& m = S.meshes[inst.geom];
    const uint32_t i0 = S.indices[m.index + h.prim * 3 + 0];
    const uint32_t i1 = S.indices[m.index + h.prim * 3 + 1];
    const uint32_t i2 = S.indices[m.index + h.prim * 3 + 2];
    const sb_vertex& v0 = S.vertices[m.vb_offset + i0];
    const sb_vertex& v1 = S.vertices[m.vb_offset + i1];
    const sb_vertex& v2 = S.vertices[m.vb_offset + i2];
    const f3 p0{ v0.pos[0], v0.pos[1], v0.pos[2] }, p1{ v1.pos[0], v1.pos[1], v1.pos[2] }, p2{ v2.pos[0], v2.pos[1], v2.pos[2] };
    const f3 n0 = unpack_normal(v0.normal), n1 = unpack_normal(v1.normal), n2 = unpack_normal(v2.normal);
    const f3 t0 = unpack_normal(v0.tangent), t1 = unpack_normal(v1.tangent), t2 = unpack_normal(v2.tangent);
    Surface s;
    s.position = xform_point(inst.o2w, interp3(p0, p1, p2, h.u, h.v)); // quirk Q12
    const f3 objN = interp3(n0, n1, n2, h.u, h.v);
    s.normal = normalize(xform_normal(inst.w2o, objN));
    s.geomNormal = normalize(xform_normal(inst.w2o, cross(p1 - p0, p2 - p0)));
    s.tangent = normalize(xform_normal(inst.w2o, interp3(t0, t1, t2, h.u, h.v))); // quirk Q11
    {
        // interpolateAttrib(uv0, uv1, uv2, barycentrics), closest_hit.cu:391-396
        const f2 a = unpack_uv(v0.uv), b = unpack_uv(v1.uv), c = unpack_uv(v2.uv);
        const float w0 = 1.0f - h.u - h.v;
        s.uv = f2{ a.x * w0 + b.x * h.u + c.x * h.v, a.y * w0 + b.y * h.u + c.y * h.v };
    }
    const float flip = inside ? -1.0f : 1.0f;
    s.geomNormal *= flip;
    s.normal *= flip;
    return s;
}

// fillCurveGeomData, closest_hit.cu:423-454
Surface curve_surface(const orc_scene& S, const InstanceData& inst, const Hit& h, const f3& rayO, const f3& rayD, bool inside)
{
    // recover the first control point index of segment h.prim (segmentIndices[prim])
    const sb_curve& c = S.curves[inst.geom];
    uint32_t offsetInside = 0, prim = 0, first = 0;
    for (uint32_t ci = 0; ci < c.vertex_counts_count; ++ci)
    {
        const uint32_t ncp = S.curveVertexCounts[c.vertex_counts_start + ci];
        const uint32_t nseg = ncp >= 3 ? ncp - 3 : 0;
        if (h.prim < prim + nseg)
        {
            first = c.points_start + offsetInside + (h.prim - prim);
            break;
        }
        prim += nseg;
        offsetInside += ncp;
    }
    f4 q[4];
    for (int k = 0; k < 4; ++k)
        q[k] = mk4(S.curvePoints[first + k], first + k < S.curveRadii.size() ? S.curveRadii[first + k] : 0.0f);
    const CubicSeg bc = cubic_from_bspline(q);
    f3 hitPoint = rayO + h.t * rayD; // getHitPoint, closest_hit.cu:327-334
    hitPoint = xform_point(inst.w2o, hitPoint);
    Surface s;
    s.normal = normalize(xform_normal(inst.w2o, cubic_surface_normal(bc, h.u, hitPoint)));
    s.tangent = normalize(xform_normal(inst.w2o, cubic_tangent(bc, h.u)));
    s.normal *= (inside ? -1.0f : 1.0f);
    s.position = xform_point(inst.o2w, hitPoint);
    s.geomNormal = s.normal; // quirk Q14
    return s;
}

// __closesthit__radiance, closest_hit.cu:456-606
void hit_surface(const RenderParams& P, const InstanceData& inst, const Hit& h, const f3& rayO, const f3& rayD, Prd& prd,
                 Counters& cnt)
{
    const orc_scene& S = *P.scene;
    const bool isInside = prd.inside;
    const Surface sf0 = (h.kind == 1) ? tri_surface(S, inst, h, isInside) : curve_surface(S, inst, h, rayO, rayD, isInside);
    const sb_material& matRecord = S.materials[inst.material];
    // mdlcode_init (closest_hit.cu:502) evaluates the material's texture inputs for this hit: the UsdUVTexture read by
    // diffuseColor replaces the constant, the one read by `normal` (scale 2, bias -1) perturbs state.normal in the
    // tangent frame (tangent_u, tangent_v = cross(N, T), N) (closest_hit.cu:408, 484-486)
    sb_material mat = matRecord;
    Surface sfm = sf0;
    if (h.kind == 1 && !S.textures.empty())
    {
        if (mat.diffuse_texture >= 1 && mat.diffuse_texture <= S.textures.size())
        {
            const f4 c = texture_lookup(S.textures[mat.diffuse_texture - 1], sf0.uv.x, sf0.uv.y);
            mat.base_color[0] = c.x;
            mat.base_color[1] = c.y;
            mat.base_color[2] = c.z;
        }
        if (mat.normal_texture >= 1 && mat.normal_texture <= S.textures.size())
        {
            const f4 c = texture_lookup(S.textures[mat.normal_texture - 1], sf0.uv.x, sf0.uv.y);
            const f3 nts{ c.x * 2.0f - 1.0f, c.y * 2.0f - 1.0f, c.z * 2.0f - 1.0f };
            const f3 bitangent = cross(sf0.normal, sf0.tangent);
            const f3 nw = nts.x * sf0.tangent + nts.y * bitangent + nts.z * sf0.normal;
            if (dot(nw, nw) > 0.0f)
                sfm.normal = normalize(nw);
        }
    }
    const Surface& sf = sfm;

    if (P.st.debug == 1)
    {
        prd.radiance = (sf.normal + mk3(1.0f)) * 0.5f; // closest_hit.cu:504-508
        return;
    }
    const f4 xi{ rnd(kBSDF0, prd.sampler), rnd(kBSDF1, prd.sampler), rnd(kBSDF2, prd.sampler), rnd(kBSDF3, prd.sampler) };
    const f3 k1 = -rayD;
    const BsdfSample bs = bsdf_sample(mat, sf.normal, sf.geomNormal, sf.tangent, k1, xi);
    if (bs.event == EV_ABSORB)
    {
        if (prd.depth == 0)
            prd.firstEventType = 1;
        prd.throughput = mk3(0.0f);
        return;
    }
    prd.specularBounce = (bs.event & EV_SPECULAR) != 0;
    if (prd.depth == 0)
    {
        if (bs.event & EV_DIFFUSE)
            prd.firstEventType = 2;
        if (bs.event & EV_GLOSSY)
            prd.firstEventType = 3;
    }
    if (bs.event & (EV_DIFFUSE | EV_GLOSSY))
    {
        // estimateDirectLighting + sampleLight, closest_hit.cu:260-324
        f3 toLight = mk3(0.0f);
        float lightPdf = 0.0f;
        f3 radiance = mk3(0.0f);
        const uint32_t numLights = uint32_t(S.lights.size());
        if (numLights > 0) // numLights == 0 is undefined in the reference (quirk Q18): no NEE here
        {
            const float u = rnd(kLightId, prd.sampler);
            uint32_t lightId = uint32_t(float(numLights) * u);
            if (lightId >= numLights)
                lightId = numLights - 1; // unreachable (u < 1); guards the array access only
            const float lightSelectionPdf = 1.0f / float(numLights);
            const sb_light& l = S.lights[lightId];
            const float ux = rnd(kLightPointX, prd.sampler), uy = rnd(kLightPointY, prd.sampler);
            const LightSample ls = sample_light(l, ux, uy, sf.position, P.st.rect_light_sampling_method);
            toLight = ls.L;
            const f3 Li{ l.color[0], l.color[1], l.color[2] };
            if (dot(sf.normal, ls.L) > 0.0f && -dot(ls.L, ls.normal) > 0.0 && all_nonzero(Li))
            {
                ++cnt.shadowRays;
                const bool occluded = trace_any(S, offset_ray(sf.position, sf.geomNormal), ls.L, P.st.shadow_ray_tmin,
                                                ls.distToLight, kRayMaskShadow);
                const float vis = occluded ? 0.0f : 1.0f;
                lightPdf = ls.pdf;
                radiance = vis * Li * saturate(dot(sf.normal, ls.L));
            }
            lightPdf *= lightSelectionPdf;
        }
        if (isnan3(radiance) || std::isnan(lightPdf))
        {
            prd.radiance = f3{ 10000.0f, 0.0f, 0.0f }; // quirk Q17
            prd.throughput = mk3(0.0f);
            return;
        }
        const bool nextEventValid = ((dot(toLight, sf.normal) > 0.0f) != isInside) && lightPdf != 0.0f;
        if (nextEventValid)
        {
            const BsdfEval ev = bsdf_evaluate(mat, sf.normal, sf.geomNormal, sf.tangent, k1, toLight);
            if (isnan3(ev.diffuse) || isnan3(ev.glossy))
            {
                prd.radiance = f3{ 10000.0f, 0.0f, 0.0f };
                prd.throughput = mk3(0.0f);
                return;
            }
            if (ev.pdf > 0.0f)
            {
                const f3 radianceOverPdf = radiance / lightPdf;
                const float w = mis_balance(lightPdf, ev.pdf);
                prd.radiance += prd.throughput * radianceOverPdf * w * (ev.diffuse + ev.glossy);
            }
        }
    }
    if (bs.event & EV_TRANSMISSION)
    {
        prd.inside = !prd.inside;
        prd.origin = offset_ray(sf.position, -sf.geomNormal);
    }
    else
    {
        prd.origin = offset_ray(sf.position, sf.geomNormal);
    }
    prd.lastBsdfPdf = prd.specularBounce ? 1.0f : bs.pdf; // quirk Q19
    prd.dir = bs.k2;
    prd.throughput *= bs.bsdf_over_pdf;
}

// one path == one iteration of the sample loop of __raygen__rg, OptixRender.cu:94-167
f3 trace_path(const RenderParams& P, uint32_t px, uint32_t py, uint32_t sampleIndex, Counters& cnt, uint8_t* firstEvent)
{
    const orc_scene& S = *P.scene;
    Prd prd{};
    prd.sampler = init_sampler(px, py, sampleIndex, P.st.spp_total, 52u);
    prd.radiance = mk3(0.0f);
    prd.throughput = mk3(1.0f);
    prd.inside = false;
    prd.depth = 0;
    prd.specularBounce = false;
    prd.lastBsdfPdf = 0.0f;
    prd.firstEventType = 0;
    f3 o, d;
    camera_ray(P, px, py, prd.sampler, o, d);
    prd.origin = o;
    prd.dir = d;
    ++cnt.paths;
    while (prd.depth < P.st.depth)
    {
        ++cnt.radianceRays;
        const Hit h = trace_closest(S, o, d, P.st.material_ray_tmin, 1e16f, kRayMaskPrimary);
        if (h.kind == 0)
        {
            // __miss__ms, OptixRender.cu:250-257 (bg_color = 0, OptixRender.cpp:739)
            prd.radiance += prd.throughput * mk3(0.0f);
            prd.throughput = mk3(0.0f);
            prd.depth = P.st.depth;
        }
        else
        {
            const InstanceData& inst = S.instances[h.inst];
            if (inst.type == SB_INSTANCE_LIGHT)
                hit_light(P, inst, o, d, h.t, prd);
            else
                hit_surface(P, inst, h, o, d, prd, cnt);
        }
        o = prd.origin;
        d = prd.dir;
        if (prd.depth > 3)
        {
            const float p = maxcomp(prd.throughput);
            if (rnd(kRussianRoulette, prd.sampler) > p)
                break;
            prd.throughput *= 1.0f / (p + 1e-5f);
        }
        if (dot(prd.throughput, prd.throughput) < 1e-5f)
            break;
        ++prd.depth;
        if (P.st.debug == 1)
            break;
        prd.sampler.depth++;
    }
    if (firstEvent)
        *firstEvent = prd.firstEventType;
    return prd.radiance;
}

// accumulate(), OptixRender.cu:60-78
f3 ref_accumulate(const f3& prev, const f3& value, const f3& exposure, uint32_t subFrameIndex)
{
    f3 c = value;
    if (subFrameIndex > 0)
    {
        const float a = 1.0f / float(subFrameIndex + 1);
        c = inverse_tonemap3(lerp(tonemap3(prev, exposure), tonemap3(c, exposure), a), exposure);
    }
    return c;
}

// exposure, OptixRender.cpp:956-987
f3 compute_exposure(const sb_settings& st)
{
    f3 e = mk3(1.0f) / mk3(1.0f);
    const float lum = e.x * 0.299f + e.y * 0.587f + e.z * 0.114f; // sutil dot, host code, unfused
    if (st.film_iso > 0.0f)
        e *= st.cm2_factor * st.film_iso / (st.shutter_speed * st.f_stop * st.f_stop) / 100.0f;
    else
        e *= st.cm2_factor;
    const float inv = 1.0f / lum; // operator/=(float3,float), sutil/vec_math.h:496-500
    e *= inv;
    return e;
}

} // namespace

extern "C" {

orc_scene* orc_scene_create(const sb_scene_view* view, uint32_t curveSplit)
{
    orc_scene* s = new orc_scene();
    s->curveSplit = curveSplit ? curveSplit : 8u;
    build_scene(*s, *view);
    return s;
}

void orc_scene_destroy(orc_scene* s)
{
    delete s;
}

void orc_scene_info(const orc_scene* s, uint64_t* out /* [4]: tris, segs, triNodes, segNodes */)
{
    out[0] = s->tris.size();
    out[1] = s->segs.size();
    out[2] = s->triBvh.nodes.size();
    out[3] = s->segBvh.nodes.size();
}

// Emulates `launches` consecutive OptiXRender::render() calls (OptixRender.cpp:989-1043) starting at
// subframe index `subframe` on the persistent `accum` buffer (float4 per pixel, may be NULL when
// subframe == 0 and the caller does not care).  `image` receives params.image (before the optional
// post-process tonemap).  counters: [paths, radiance rays, shadow rays].  Returns the new subframe.
uint32_t orc_render(const orc_scene* scene, const sb_settings* st, const float* clipToView, const float* viewToWorld,
                    uint32_t width, uint32_t height, uint32_t subframe, uint32_t launches, float* accum, float* image,
                    uint64_t* counters, int nthreads, float* aov)
{
    RenderParams P;
    P.scene = scene;
    P.st = *st;
    std::memcpy(P.clipToView, clipToView, sizeof(P.clipToView));
    std::memcpy(P.viewToWorld, viewToWorld, sizeof(P.viewToWorld));
    P.width = width;
    P.height = height;
    P.exposure = compute_exposure(*st);
    if (nthreads <= 0)
        nthreads = int(std::thread::hardware_concurrency());
    if (nthreads <= 0)
        nthreads = 1;

    // per-launch sample counts (OptixRender.cpp:989-1000)
    std::vector<uint32_t> launchSamples, launchStart;
    uint32_t sub = subframe;
    const bool debugNormals = (st->debug == 1);
    bool acc = st->enable_acc != 0;
    for (uint32_t l = 0; l < launches; ++l)
    {
        const int32_t left = int32_t(st->spp_total) - int32_t(sub);
        uint32_t n = acc ? uint32_t(std::max(0, std::min(int32_t(st->spp), left))) : st->spp;
        if (debugNormals)
        {
            n = 1;
            acc = false;
        }
        launchSamples.push_back(n);
        launchStart.push_back(sub);
        if (n != 0)
            sub = acc ? sub + n : 0;
    }

    std::atomic<uint32_t> nextRow{ 0 };
    std::vector<Counters> cnts(nthreads);
    auto worker = [&](int tid) {
        Counters& cnt = cnts[tid];
        for (;;)
        {
            const uint32_t y = nextRow.fetch_add(1);
            if (y >= height)
                break;
            for (uint32_t x = 0; x < width; ++x)
            {
                const size_t pix = size_t(y) * width + x;
                f3 accumC = accum ? f3{ accum[pix * 4], accum[pix * 4 + 1], accum[pix * 4 + 2] } : mk3(0.0f);
                f3 img = accumC;
                // AOV state of this pixel: params.diffuse / specular + their uint16 counters (OptixRender.cu:169-221)
                f3 aovD = aov ? f3{ aov[pix * 10], aov[pix * 10 + 1], aov[pix * 10 + 2] } : mk3(0.0f);
                f3 aovS = aov ? f3{ aov[pix * 10 + 4], aov[pix * 10 + 5], aov[pix * 10 + 6] } : mk3(0.0f);
                uint32_t cntD = aov ? uint32_t(aov[pix * 10 + 8]) : 0u, cntS = aov ? uint32_t(aov[pix * 10 + 9]) : 0u;
                for (size_t l = 0; l < launchSamples.size(); ++l)
                {
                    const uint32_t n = launchSamples[l];
                    if (n == 0)
                    {
                        // copy of accum / diffuse / specular to image, OptixRender.cpp:1020-1043
                        img = st->debug == 2 ? aovD : (st->debug == 3 ? aovS : accumC);
                        continue;
                    }
                    f3 result = mk3(0.0f), diffuse = mk3(0.0f), specular = mk3(0.0f);
                    uint32_t nD = 0, nS = 0;
                    for (uint32_t sidx = 0; sidx < n; ++sidx)
                    {
                        uint8_t ev = 0;
                        const f3 r = trace_path(P, x, y, launchStart[l] + sidx, cnt, &ev);
                        result += r;
                        if (ev == 2)
                        {
                            diffuse += r;
                            ++nD;
                        }
                        if (ev == 3)
                        {
                            specular += r;
                            ++nS;
                        }
                    }
                    result = result / float(n);
                    const uint32_t sub0 = launchStart[l];
                    f3 diffuseOut, specularOut;
                    if (nD > 0)
                    {
                        diffuse = diffuse / float(nD);
                        const uint32_t prev = sub0 > 0 ? cntD : 0u;
                        aovD = ref_accumulate(aovD, diffuse, P.exposure, prev);
                        diffuseOut = aovD;
                        cntD = uint16_t(prev + nD);
                    }
                    else
                    {
                        if (sub0 == 0)
                        {
                            aovD = mk3(0.0f);
                            cntD = 0;
                        }
                        diffuseOut = aovD;
                    }
                    if (nS > 0)
                    {
                        specular = specular / float(nS);
                        const uint32_t prev = sub0 > 0 ? cntS : 0u;
                        aovS = ref_accumulate(aovS, specular, P.exposure, prev);
                        specularOut = aovS;
                        cntS = uint16_t(prev + nS);
                    }
                    else
                    {
                        if (sub0 == 0)
                        {
                            aovS = mk3(0.0f);
                            cntS = 0;
                        }
                        specularOut = cntS > 0 ? aovS : mk3(0.0f);
                    }
                    if (st->debug == 2)
                    {
                        img = diffuseOut;
                        continue;
                    }
                    if (st->debug == 3)
                    {
                        img = specularOut;
                        continue;
                    }
                    if (acc && st->debug == 0)
                    {
                        accumC = ref_accumulate(accumC, result, P.exposure, launchStart[l]);
                        img = accumC;
                    }
                    else
                    {
                        img = result;
                    }
                }
                if (accum)
                {
                    accum[pix * 4] = accumC.x;
                    accum[pix * 4 + 1] = accumC.y;
                    accum[pix * 4 + 2] = accumC.z;
                    accum[pix * 4 + 3] = 1.0f;
                }
                if (aov)
                {
                    const float a[10] = { aovD.x, aovD.y, aovD.z, 1.0f, aovS.x, aovS.y, aovS.z, 1.0f, float(cntD), float(cntS) };
                    std::memcpy(aov + pix * 10, a, sizeof(a));
                }
                image[pix * 4] = img.x;
                image[pix * 4 + 1] = img.y;
                image[pix * 4 + 2] = img.z;
                image[pix * 4 + 3] = 1.0f;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t)
        pool.emplace_back(worker, t);
    worker(0);
    for (auto& t : pool)
        t.join();
    if (counters)
    {
        for (const Counters& c : cnts)
        {
            counters[0] += c.paths;
            counters[1] += c.radianceRays;
            counters[2] += c.shadowRays;
        }
    }
    return sub;
}

// Individual radiance samples (before accumulation) for a list of (x, y, sampleIndex): 3 floats each.
// nthreads <= 0: all host threads.  counters (optional) [paths, radiance rays, shadow rays] are added to.
void orc_path_radiance(const orc_scene* scene, const sb_settings* st, const float* clipToView, const float* viewToWorld,
                       uint32_t width, uint32_t height, uint32_t n, const uint32_t* xs, const uint32_t* ys,
                       const uint32_t* samples, float* out, int nthreads, uint64_t* counters)
{
    RenderParams P;
    P.scene = scene;
    P.st = *st;
    std::memcpy(P.clipToView, clipToView, sizeof(P.clipToView));
    std::memcpy(P.viewToWorld, viewToWorld, sizeof(P.viewToWorld));
    P.width = width;
    P.height = height;
    P.exposure = compute_exposure(*st);
    if (nthreads <= 0)
        nthreads = int(std::thread::hardware_concurrency());
    if (nthreads <= 0)
        nthreads = 1;
    std::vector<Counters> cnts(nthreads);
    std::atomic<uint32_t> next{ 0 };
    const uint32_t kGrab = 256;
    auto worker = [&](int tid) {
        for (;;)
        {
            const uint32_t b = next.fetch_add(kGrab);
            if (b >= n)
                break;
            for (uint32_t i = b; i < std::min(n, b + kGrab); ++i)
            {
                const f3 r = trace_path(P, xs[i], ys[i], samples[i], cnts[tid], nullptr);
                out[3 * size_t(i)] = r.x;
                out[3 * size_t(i) + 1] = r.y;
                out[3 * size_t(i) + 2] = r.z;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t)
        pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool)
        th.join();
    if (counters)
        for (const Counters& c : cnts)
        {
            counters[0] += c.paths;
            counters[1] += c.radianceRays;
            counters[2] += c.shadowRays;
        }
}

void orc_exposure(const sb_settings* st, float* out3)
{
    const f3 e = compute_exposure(*st);
    out3[0] = e.x;
    out3[1] = e.y;
    out3[2] = e.z;
}

// rays: 8 floats (o, tmin, d, tmax); mode 0 closest / mask 255, mode 1 any / mask 3
void orc_trace(const orc_scene* scene, uint32_t n, const float* rays, uint32_t mode, sb_hit* hits)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const float* r = rays + 8 * i;
        const f3 o{ r[0], r[1], r[2] }, d{ r[4], r[5], r[6] };
        if (mode == 0)
        {
            const Hit h = trace_closest(*scene, o, d, r[3], r[7], kRayMaskPrimary);
            hits[i] = sb_hit{ h.t, h.u, h.v, h.prim, h.inst, h.kind };
        }
        else
        {
            const bool occ = trace_any(*scene, o, d, r[3], r[7], kRayMaskShadow);
            hits[i] = sb_hit{ 0, 0, 0, 0, 0, occ ? 1u : 0u };
        }
    }
}

// ---- function-level hooks for the golden-vector tests -------------------------------------------

void orc_sampler(uint32_t n, const uint32_t* x, const uint32_t* y, const uint32_t* sample, const uint32_t* maxSamples,
                 const uint32_t* depth, const uint32_t* dim, float* out)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        Sampler s = init_sampler(x[i], y[i], sample[i], maxSamples[i], 52u);
        s.depth = depth[i];
        out[i] = rnd(SampleDim(dim[i]), s);
    }
}

void orc_sampler_ints(uint32_t* out /* [7] */)
{
    out[0] = morton2(3, 5);
    out[1] = morton2(1023, 767);
    out[2] = fmix32(52);
    out[3] = seed_combine(fmix32(52), 3);
    out[4] = lk_permute(1, 2);
    out[5] = owen_scramble(12345, fmix32(52));
    out[6] = sobol_u32(5, 2);
}

void orc_sobol_table(uint32_t* out /* [5*32] */)
{
    std::memcpy(out, sobol_table().v, sizeof(uint32_t) * 160);
}

void orc_light_sample(uint32_t n, const sb_light* lights, const float* hitPoints, const float* u, uint32_t method, float* out)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const f3 hp{ hitPoints[3 * i], hitPoints[3 * i + 1], hitPoints[3 * i + 2] };
        const LightSample s = sample_light(lights[i], u[2 * i], u[2 * i + 1], hp, method);
        float* o = out + 12 * i;
        o[0] = s.pointOnLight.x;
        o[1] = s.pointOnLight.y;
        o[2] = s.pointOnLight.z;
        o[3] = s.pdf;
        o[4] = s.normal.x;
        o[5] = s.normal.y;
        o[6] = s.normal.z;
        o[7] = s.area;
        o[8] = s.L.x;
        o[9] = s.L.y;
        o[10] = s.L.z;
        o[11] = s.distToLight;
    }
}

// getLightPdf(l, lightHit, surfaceHit) and calcLightNormal: out 4 floats (pdf, n.xyz)
void orc_light_pdf(uint32_t n, const sb_light* lights, const float* lightHits, const float* surfaceHits, float* out)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const f3 lh{ lightHits[3 * i], lightHits[3 * i + 1], lightHits[3 * i + 2] };
        const f3 sh{ surfaceHits[3 * i], surfaceHits[3 * i + 1], surfaceHits[3 * i + 2] };
        const f3 nn = light_normal(lights[i], lh);
        out[4 * i] = light_pdf(lights[i], lh, sh);
        out[4 * i + 1] = nn.x;
        out[4 * i + 2] = nn.y;
        out[4 * i + 3] = nn.z;
    }
}

float orc_mis_balance(float a, float b)
{
    return mis_balance(a, b);
}

// mode 0: tonemap, 1: inverseTonemap, 2: accumulate(prev=c, value=c2, subframe)
void orc_tonemap(uint32_t mode, const float* c, const float* c2, const float* exposure, uint32_t subframe, float* out)
{
    const f3 a{ c[0], c[1], c[2] }, e{ exposure[0], exposure[1], exposure[2] };
    f3 r;
    if (mode == 0)
        r = tonemap3(a, e);
    else if (mode == 1)
        r = inverse_tonemap3(a, e);
    else
        r = ref_accumulate(a, f3{ c2[0], c2[1], c2[2] }, e, subframe);
    out[0] = r.x;
    out[1] = r.y;
    out[2] = r.z;
}

// q: 16 floats (4 control points xyzr); out: pos4(u) [4], vel4(u) [4], tangent [3], normal [3], ps' [3]
void orc_curve_eval(const float* q, float u, const float* ps, float* out)
{
    f4 cp[4];
    for (int k = 0; k < 4; ++k)
        cp[k] = f4{ q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3] };
    const CubicSeg bc = cubic_from_bspline(cp);
    const f4 p = cubic_position4(bc, u), v = cubic_velocity4(bc, u);
    const f3 tg = cubic_tangent(bc, u);
    f3 pp{ ps[0], ps[1], ps[2] };
    const f3 nn = cubic_surface_normal(bc, u, pp);
    const float r[17] = { p.x, p.y, p.z, p.w, v.x, v.y, v.z, v.w, tg.x, tg.y, tg.z, nn.x, nn.y, nn.z, pp.x, pp.y, pp.z };
    std::memcpy(out, r, sizeof(r));
}

// the float solver used by renders; out: hit, t, u
void orc_curve_intersect_f32(const float* q, const float* ray, float* out)
{
    f4 cp[4];
    for (int k = 0; k < 4; ++k)
        cp[k] = f4{ q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3] };
    float t = 0.0f, u = 0.0f;
    const CurveSpan sp = curve_span(cp, 0, 1); // the whole segment as one span
    const bool hit = intersect_round_cubic_f32(sp.c, f3{ ray[0], ray[1], ray[2] }, f3{ ray[3], ray[4], ray[5] }, ray[6], ray[7], t, u);
    out[0] = hit ? 1.0f : 0.0f;
    out[1] = t;
    out[2] = u;
}

// the double-precision bracketing solver (validator); ray: o[3], d[3], tmin, tmax; out: hit, t, u
void orc_curve_intersect(const float* q, const float* ray, float* out)
{
    f4 cp[4];
    for (int k = 0; k < 4; ++k)
        cp[k] = f4{ q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3] };
    const CurveHit h = intersect_round_cubic(cp, f3{ ray[0], ray[1], ray[2] }, f3{ ray[3], ray[4], ray[5] }, ray[6], ray[7]);
    out[0] = h.hit ? 1.0f : 0.0f;
    out[1] = h.t;
    out[2] = h.u;
}

void orc_offset_ray(const float* p, const float* n, float* out)
{
    const f3 r = offset_ray(f3{ p[0], p[1], p[2] }, f3{ n[0], n[1], n[2] });
    out[0] = r.x;
    out[1] = r.y;
    out[2] = r.z;
}

// texture lookup hook: n (u, v) pairs on texture `index0` (0-based) -> rgba
void orc_texture_lookup(const orc_scene* scene, uint32_t index0, uint32_t n, const float* uv, float* out)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const f4 c = texture_lookup(scene->textures[index0], uv[2 * i], uv[2 * i + 1]);
        out[4 * i] = c.x;
        out[4 * i + 1] = c.y;
        out[4 * i + 2] = c.z;
        out[4 * i + 3] = c.w;
    }
}

// BSDF hook, batched.  in: 19 floats per item (n[3], ng[3], tangent[3], k1[3], xi[4], k2 for evaluate[3]);
// out: 15 floats per item (sample: k2[3], bsdf_over_pdf[3], pdf, event; evaluate: diffuse[3], glossy[3], pdf)
void orc_bsdf_batch(const sb_material* m, uint32_t n, const float* in, float* out)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const float* a = in + 19 * size_t(i);
        float* o = out + 15 * size_t(i);
        const f3 N{ a[0], a[1], a[2] }, NG{ a[3], a[4], a[5] }, T{ a[6], a[7], a[8] }, K1{ a[9], a[10], a[11] };
        const BsdfSample s = bsdf_sample(*m, N, NG, T, K1, f4{ a[12], a[13], a[14], a[15] });
        o[0] = s.k2.x;
        o[1] = s.k2.y;
        o[2] = s.k2.z;
        o[3] = s.bsdf_over_pdf.x;
        o[4] = s.bsdf_over_pdf.y;
        o[5] = s.bsdf_over_pdf.z;
        o[6] = s.pdf;
        o[7] = float(s.event);
        const BsdfEval e = bsdf_evaluate(*m, N, NG, T, K1, f3{ a[16], a[17], a[18] });
        o[8] = e.diffuse.x;
        o[9] = e.diffuse.y;
        o[10] = e.diffuse.z;
        o[11] = e.glossy.x;
        o[12] = e.glossy.y;
        o[13] = e.glossy.z;
        o[14] = e.pdf;
    }
}

// Camera matrices as uploaded by the reference (OptixRender.cpp:895-897, 953-954; camera.cpp:61-131):
// view = glm column-major view matrix; outputs are row-major Params.clipToView / viewToWorld.
void orc_camera_matrices(const float* view, float fovYDeg, float aspect, float* clipToView, float* viewToWorld)
{
    // perspective(): focal_length = 1/tan(radians(fov)/2); x = f/aspect; y = f.  Reverse-Z near/far
    // only enter rows 2,3 which a clip vector (x,y,1,1) maps to view z = -1 and an unused w.
    const float focal = 1.0f / std::tan((fovYDeg * 0.01745329251994329576923690768489f) / 2.0f);
    const float x = focal / aspect, y = focal;
    for (int i = 0; i < 16; ++i)
        clipToView[i] = 0.0f;
    clipToView[0] = 1 / x;
    clipToView[5] = 1 / y;
    clipToView[11] = -1.0f;
    // row 3 = (0, 0, 1/B, A/B) depends on near/far; it only produces the unused w component
    clipToView[14] = 0.0f;
    clipToView[15] = 1.0f;
    // viewToWorld = rowMajor(inverse(view)); view is rigid+scale affine in practice -> general inverse in double
    double m[16], inv[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
            m[r * 4 + c] = view[c * 4 + r]; // to row-major
    // Gauss-Jordan
    double a[4][8];
    for (int r = 0; r < 4; ++r)
    {
        for (int c = 0; c < 4; ++c)
        {
            a[r][c] = m[r * 4 + c];
            a[r][4 + c] = (r == c) ? 1.0 : 0.0;
        }
    }
    for (int col = 0; col < 4; ++col)
    {
        int piv = col;
        for (int r = col + 1; r < 4; ++r)
            if (std::fabs(a[r][col]) > std::fabs(a[piv][col]))
                piv = r;
        for (int c = 0; c < 8; ++c)
            std::swap(a[col][c], a[piv][c]);
        const double d = 1.0 / a[col][col];
        for (int c = 0; c < 8; ++c)
            a[col][c] *= d;
        for (int r = 0; r < 4; ++r)
        {
            if (r == col)
                continue;
            const double f = a[r][col];
            for (int c = 0; c < 8; ++c)
                a[r][c] -= f * a[col][c];
        }
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
            inv[r * 4 + c] = a[r][4 + c];
    for (int i = 0; i < 16; ++i)
        viewToWorld[i] = float(inv[i]);
}

// Post-process of OptiXRender::render (OptixRender.cpp:1045-1049; Tonemappers.cu:17-135), in place on
// float4 pixels.  type: 0 none, 1 Reinhard, 2 ACES fitted, 3 ACES film.
void orc_postprocess(float* image, uint32_t npix, uint32_t type, const float* exposure, float gamma)
{
    const f3 e{ exposure[0], exposure[1], exposure[2] };
    for (uint32_t i = 0; i < npix; ++i)
    {
        f3 c{ image[4 * i], image[4 * i + 1], image[4 * i + 2] };
        if (type == 1)
        {
            const f3 r = c * e;
            const float lum = r.x * 0.299f + r.y * 0.587f + r.z * 0.114f;
            c = r / (lum + 1);
        }
        else if (type == 2)
        {
            const f3 r = c * e;
            // ACESInputMat * color (double literals rounded to float by sutil::Matrix3x3)
            f3 v{ 0.59719f * r.x + 0.35458f * r.y + 0.04823f * r.z, 0.07600f * r.x + 0.90834f * r.y + 0.01566f * r.z,
                  0.02840f * r.x + 0.13383f * r.y + 0.83777f * r.z };
            const f3 a = v * (v + mk3(0.0245786f)) - mk3(0.000090537f);
            const f3 b = v * (0.983729f * v + mk3(0.4329510f)) + mk3(0.238081f);
            v = a / b;
            c = f3{ 1.60475f * v.x + -0.53108f * v.y + -0.07367f * v.z, -0.10208f * v.x + 1.10813f * v.y + -0.00605f * v.z,
                    -0.00327f * v.x + -0.07276f * v.y + 1.07602f * v.z };
            c = f3{ saturate(c.x), saturate(c.y), saturate(c.z) };
        }
        else if (type == 3)
        {
            const f3 x = c * e;
            const float A = 2.51f, B = 0.03f, C = 2.43f, D = 0.59f, E = 0.14f;
            const f3 r = (x * (A * x + mk3(B))) / (x * (C * x + mk3(D)) + mk3(E));
            c = f3{ saturate(r.x), saturate(r.y), saturate(r.z) };
        }
        if (gamma > 0.0f)
        {
            const float ig = 1.0f / gamma;
            c = f3{ std::pow(c.x, ig), std::pow(c.y, ig), std::pow(c.z, ig) };
        }
        if (type != 0 || gamma > 0.0f)
        {
            image[4 * i] = c.x;
            image[4 * i + 1] = c.y;
            image[4 * i + 2] = c.z;
            image[4 * i + 3] = 1.0f;
        }
    }
}

} // extern "C"
