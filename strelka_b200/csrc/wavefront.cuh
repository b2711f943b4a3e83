// Per-path logic of the wavefront integrator (SB_HD: compiled by nvcc for the kernels in kernels.cu and,
// for the CPU test-suite only, by g++ in tests/emul).
//
// One camera path = one iteration of the sample loop of __raygen__rg (OptixRender.cu:94-167).  The
// reference runs it as a megakernel thread per pixel; here every bounce is three data-parallel stages
// over SoA queues:   extend (closest hit)  ->  shade (this file)  ->  shadow (any hit + NEE add)
// Surviving paths are compacted into the next queue with warp-aggregated atomics.
//
//   raygen_one : initSampler + generateCameraRay (OptixRender.cu:38-58, 96-113)
//   shade_one  : __miss__ms (:250-257), __closesthit__light (:315-341), __closesthit__radiance
//                (closest_hit.cu:456-606) and the tail of the bounce loop (OptixRender.cu:131-153)
//   shadow_one : traceOcclusion (closest_hit.cu:185-197) + the NEE add (:586)
//   accumulate : accumulate() (OptixRender.cu:60-78) in its shardable form S = sum_k T(L_k)
#pragma once
#include "sampler.cuh"
#include "lights.cuh"
#include "bsdf.cuh"
#include "bvh_build.h"

namespace sb
{

constexpr uint32_t kMaxDepth = 32;
// Device-resident queue counters (Queues::counts).  The path-queue count of bounce d+1 and the shadow-queue count
// of bounce d share one aligned 64-bit word, so that the shade kernel reserves both with ONE atomic per warp:
//   counts[2d] = rays queued for bounce d,  counts[2d + 3] = shadow rays cast at bounce d
SB_HD constexpr uint32_t count_path(uint32_t depth)
{
    return 2u * depth;
}
SB_HD constexpr uint32_t count_shadow(uint32_t depth)
{
    return 2u * depth + 3u;
}
constexpr uint32_t kHeadExtendBase = 72; // dynamic-fetch cursors of k_extend / k_shadow, one per bounce
constexpr uint32_t kHeadShadowBase = 104;
constexpr uint32_t kHeadFused = 136; // path-id dispenser of the fused small-scene kernel
constexpr uint32_t kNumCounts = 144;

// per-path flag bits (stored in thr.w)
constexpr uint32_t kFlagInside = 1u, kFlagSpecular = 2u;
constexpr uint32_t kFlagEventShift = 2u; // EventType of the first bounce (2 bits): 0 undef, 1 absorb, 2 diffuse, 3 specular

struct FrameParams
{
    uint32_t width, height, tilesX, nPixPadded;
    uint32_t maxDepth, sppTotal, rectMethod, debug;
    float shadowTmin, materialTmin;
    float clipToView[16], viewToWorld[16];
    float exposure[3];
    uint32_t sampleBase, sampleStride, chunk, numLights;
};

struct StatCounters // device-resident, persistent across launches
{
    unsigned long long paths, radianceRays, shadowRays, nodes, tris, segs, overflow, nodesSh, trisSh, segsSh;
};

struct Queues
{
    float4* rayO[2]; // origin.xyz, bits(pathId)
    float4* rayD[2]; // dir.xyz, lastBsdfPdf
    float4* thr[2]; // throughput.xyz, bits(flags)
    float4* hitA; // t, u, v, bits(global triangle id | curve SegInfo index)
    uint32_t* hitB; // instance | kind << 30
    float4* Lacc; // per pathId: radiance accumulated along the path; w = bits(first bsdf event), for the AOV views
    float4* shO; // origin.xyz, tmin
    float4* shD; // dir.xyz, tmax
    float4* shC; // contribution.xyz, bits(pathId)
    uint32_t* counts; // kNumCounts
    StatCounters* stats;
    const uint32_t* sobolTab; // kSobolTabWords, byte-sliced Sobol tables (global memory copy)
};

// The ping-pong path queue `i` (bounce & 1 on the host).  The kernels always read queue 0 and write queue 1: the
// launcher swaps the pointers per bounce.  Never index Q.rayO[] with a run-time value in a kernel: the compiler
// then copies the whole parameter struct to local memory and every queue pointer costs a local load (seen in
// the SASS of k_shade: LDC/STL prologue, LDL before each queue access).
SB_HD float4* qsel(float4* const (&q)[2], int i)
{
    return i ? q[1] : q[0];
}

// Warp-aggregated slot allocation: one atomicAdd per warp instead of one per surviving lane.
SB_HD uint32_t queue_alloc(uint32_t* counter)
{
#if defined(__CUDA_ARCH__)
    const unsigned mask = __activemask();
    const int leader = __ffs(mask) - 1;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned rank = __popc(mask & ((1u << lane) - 1u));
    uint32_t base = 0;
    if (int(lane) == leader)
        base = atomicAdd(counter, uint32_t(__popc(mask)));
    base = __shfl_sync(mask, base, leader);
    return base + rank;
#else
    return (*counter)++;
#endif
}

SB_HD void path_pixel(const FrameParams& P, uint32_t pathId, uint32_t& x, uint32_t& y, uint32_t& k)
{
    k = pathId / P.nPixPadded;
    const uint32_t p = pathId - k * P.nPixPadded;
    const uint32_t tile = p >> 5, within = p & 31u;
    x = (tile % P.tilesX) * 8u + (within & 7u); // 8x4-pixel tiles: one warp = one tile of primary rays
    y = (tile / P.tilesX) * 4u + (within >> 3);
}

// row-major 4x4 * float4 as sutil::Matrix4x4 does it (sutil/Matrix.h)
SB_HD float4 mat4_mul(const float* m, const float4& v)
{
    return mk4(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w, m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * v.w,
               m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * v.w, m[12] * v.x + m[13] * v.y + m[14] * v.z + m[15] * v.w);
}

// The state one path carries from bounce to bounce (PerRayData, OptixRender.cu:85-153): queue records in the
// wavefront kernels, registers in the fused path kernel.
struct PathState
{
    float3 o, d; // current path segment
    float lastBsdfPdf;
    float3 throughput;
    uint32_t flags;
    uint32_t pathId;
    float4 L; // radiance collected so far; w = bits(first bsdf event) for the AOV views
};

// Primary path segment of one (pixel, sample): false for the padding pixels of an edge tile.
SB_HD bool raygen_state(const FrameParams& P, uint32_t pathId, PathState& ps)
{
    uint32_t x, y, k;
    path_pixel(P, pathId, x, y, k);
    if (x >= P.width || y >= P.height)
        return false;
    const uint32_t sidx = sampler_index(x, y, P.sampleBase + k * P.sampleStride, P.sppTotal);
    const uint32_t s0 = fmix32(52u);
    const uint32_t index = owen_scramble(sidx, s0);
    const float jx = u32_to_unit(owen_scramble(sobol_u32(index, 0), seed_combine(s0, 0)));
    const float jy = u32_to_unit(owen_scramble(sobol_u32(index, 1), seed_combine(s0, 1)));
    // generateCameraRay, OptixRender.cu:38-58
    const float posx = float(x) + jx, posy = float(y) + jy;
    const float ndcx = (posx / float(P.width)) * 2.0f - 1.0f;
    const float ndcy = (posy / float(P.height)) * 2.0f - 1.0f;
    const float4 vs = mat4_mul(P.clipToView, mk4(ndcx, ndcy, 1.0f, 1.0f));
    const float4 wdir = mat4_mul(P.viewToWorld, mk4(vs.x, vs.y, vs.z, 0.0f));
    ps.o = mk3(mat4_mul(P.viewToWorld, mk4(0.0f, 0.0f, 0.0f, 1.0f)));
    ps.d = normalize(mk3(wdir));
    ps.lastBsdfPdf = 0.0f;
    ps.throughput = mk3(1.0f, 1.0f, 1.0f);
    ps.flags = 0u;
    ps.pathId = pathId;
    ps.L = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    return true;
}

SB_HD void raygen_one(const FrameParams& P, const Queues& Q, uint32_t pathId)
{
    Q.Lacc[pathId] = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    PathState ps;
    if (!raygen_state(P, pathId, ps))
        return;
    const uint32_t slot = queue_alloc(&Q.counts[0]);
    Q.rayO[0][slot] = mk4(ps.o, u2f(pathId));
    Q.rayD[0][slot] = mk4(ps.d, 0.0f);
    Q.thr[0][slot] = mk4(1.0f, 1.0f, 1.0f, u2f(0u));
}

// unpackNormal, closest_hit.cu:236-244.  A packed component has 10 bits (12 for z, of which a valid packing
// uses 10), so q / 511.99999f * 2 - 1 takes 1024 distinct values: the shade kernel reads them from a
// shared-memory table built with exactly this expression (unpack_lut_fill) instead of paying an int->float
// conversion and an IEEE division per component -- same bits.
constexpr uint32_t kUnpackLutSize = 1024;
SB_HD float unpack_component(uint32_t q)
{
    return float(q) / 511.99999f * 2.0f - 1.0f;
}
SB_HD float3 unpack_normal(uint32_t val, const float* lut = nullptr)
{
    const uint32_t qz = (val & 0xfff00000u) >> 20, qy = (val & 0x000ffc00u) >> 10, qx = val & 0x000003ffu;
    float3 n;
    n.z = (lut && qz < kUnpackLutSize) ? lut[qz] : unpack_component(qz);
    n.y = lut ? lut[qy] : unpack_component(qy);
    n.x = lut ? lut[qx] : unpack_component(qx);
    return n;
}
// offset_ray, closest_hit.cu:218-233 (Ray Tracing Gems ch. 6)
SB_HD float3 offset_ray(const float3& p, const float3& n)
{
    const float origin = 1.0f / 32.0f;
    const float float_scale = 1.0f / 65536.0f;
    const float int_scale = 256.0f;
    const int ix = int(int_scale * n.x), iy = int(int_scale * n.y), iz = int(int_scale * n.z);
    const float3 pi = mk3(u2f(uint32_t(int(f2u(p.x)) + ((p.x < 0) ? -ix : ix))), u2f(uint32_t(int(f2u(p.y)) + ((p.y < 0) ? -iy : iy))),
                          u2f(uint32_t(int(f2u(p.z)) + ((p.z < 0) ? -iz : iz))));
    return mk3(fabsf(p.x) < origin ? p.x + float_scale * n.x : pi.x, fabsf(p.y) < origin ? p.y + float_scale * n.y : pi.y,
               fabsf(p.z) < origin ? p.z + float_scale * n.z : pi.z);
}
// interpolateAttrib, closest_hit.cu:199-205
SB_HD float3 interp3(const float3& a, const float3& b, const float3& c, float bx, float by)
{
    return a * (1.0f - bx - by) + b * bx + c * by;
}

struct Surface
{
    float3 position, normal, geomNormal, tangent;
};

// fillTriangleGeomData, closest_hit.cu:365-421 (quirks Q11, Q12)
// The four 16-byte quarters of a TriShade record.
struct TriShadeRegs
{
    float4 a, b, c, d;
};
SB_HD TriShadeRegs load_tri_shade(const TriShade* rec)
{
    const float4* q = reinterpret_cast<const float4*>(rec);
    TriShadeRegs r;
    r.a = q[0];
    r.b = q[1];
    r.c = q[2];
    r.d = q[3];
    return r;
}
SB_HD Surface tri_surface(const InstDev& I, const TriShadeRegs& r, float bu, float bv, bool inside, const float* lut)
{
    const float3 p0 = mk3(r.a.x, r.a.y, r.a.z), p1 = mk3(r.a.w, r.b.x, r.b.y), p2 = mk3(r.b.z, r.b.w, r.c.x);
    const uint32_t n0 = f2u(r.c.y), n1 = f2u(r.c.z), n2 = f2u(r.c.w), t0 = f2u(r.d.x), t1 = f2u(r.d.y), t2 = f2u(r.d.z);
    Surface s;
    s.position = xform_point(I.o2w, interp3(p0, p1, p2, bu, bv));
    s.normal = normalize(xform_normal(I.w2o, interp3(unpack_normal(n0, lut), unpack_normal(n1, lut), unpack_normal(n2, lut), bu, bv)));
    s.geomNormal = normalize(xform_normal(I.w2o, cross(p1 - p0, p2 - p0)));
    s.tangent = normalize(xform_normal(I.w2o, interp3(unpack_normal(t0, lut), unpack_normal(t1, lut), unpack_normal(t2, lut), bu, bv)));
    const float flip = inside ? -1.0f : 1.0f;
    s.geomNormal *= flip;
    s.normal *= flip;
    return s;
}

// fillCurveGeomData, closest_hit.cu:423-454 (quirk Q14)
SB_HD Surface curve_surface(const SceneDev& S, const InstDev& I, uint32_t segIndex, float u, float t, const float3& rayO, const float3& rayD,
                            bool inside)
{
    const uint32_t first = S.segInfo[segIndex].firstPoint;
    float4 q[4];
    for (int k = 0; k < 4; ++k)
    {
        const uint32_t pi = first + k;
        q[k] = mk4(S.curvePoints[3 * pi], S.curvePoints[3 * pi + 1], S.curvePoints[3 * pi + 2], pi < S.numCurveRadii ? S.curveRadii[pi] : 0.0f);
    }
    const CubicSeg bc = cubic_from_bspline(q);
    float3 hitPoint = rayO + t * rayD;
    hitPoint = xform_point(I.w2o, hitPoint);
    Surface s;
    s.normal = normalize(xform_normal(I.w2o, cubic_surface_normal(bc, u, hitPoint)));
    s.tangent = normalize(xform_normal(I.w2o, cubic_tangent(bc, u)));
    s.normal *= (inside ? -1.0f : 1.0f);
    s.position = xform_point(I.o2w, hitPoint);
    s.geomNormal = s.normal;
    return s;
}

// One bounce of one path after its closest-hit query (ha = t,u,v,bits(prim); hb = instance | kind << 30).
// `depth` == prd.depth == sampler.depth.  Returns true when the path continues with the updated PathState.
// Side effects go through `sink` at the point they arise (keeps the live ranges short):
//   sink.radiance_changed(ps)                      ps.L changed (emitter hit, debug view, NaN guard, AOV tag)
//   sink.shadow_ray(ps, shO, shD, contrib)         trace (origin.xyz + tmin, dir.xyz + tmax); add contrib to L if unoccluded
// CURVES / PREVIEW = false compile the curve attributes / the UsdPreviewSurface model out for scenes without them,
// RECT_UNIFORM = true everything but the uniform rect-light sampler (all lights rect, rectLightSamplingMethod 0),
// HAIR = false the hair fibre BSDF (the shade kernel is register-bound; the host picks the variant per scene)
// unpackUV, closest_hit.cu:246-254: 16-16 bit st over [-10, 10] (v already flipped by the delegate, RenderPass.cpp:109-114)
SB_HD float2 unpack_uv(uint32_t val)
{
    float2 uv;
    uv.y = float((val & 0xffff0000u) >> 16) / 16383.99999f * 20.0f - 10.0f;
    uv.x = float(val & 0x0000ffffu) / 16383.99999f * 20.0f - 10.0f;
    return uv;
}

// CUDA's documented linear filter (Programming Guide, "Texture Fetching": normalised coordinates, wrap addressing,
// xB = x N - 0.5, i = floor(xB), alpha = frac(xB) kept as 1.8 fixed point) over RGBA8 texels read as normalised floats:
// what tex2D<float4> does in hardware for the texture objects the reference creates (OptixRender.cpp:1229-1252).
// Only the host emulation of the tests runs this function; device code samples the texture object.
SB_HD float4 texture_bilinear_rgba8(const uint8_t* px, uint32_t w, uint32_t h, float u, float v)
{
    const float xb = (u - floorf(u)) * float(w) - 0.5f, yb = (v - floorf(v)) * float(h) - 0.5f;
    const float xf = floorf(xb), yf = floorf(yb);
    const float a = floorf((xb - xf) * 256.0f + 0.5f) * (1.0f / 256.0f), b = floorf((yb - yf) * 256.0f + 0.5f) * (1.0f / 256.0f);
    const int x0 = ((int(xf) % int(w)) + int(w)) % int(w), y0 = ((int(yf) % int(h)) + int(h)) % int(h);
    const int x1 = (x0 + 1) % int(w), y1 = (y0 + 1) % int(h);
    float r[4];
    for (int c = 0; c < 4; ++c)
    {
        const float t00 = float(px[4 * (size_t(y0) * w + x0) + c]) * (1.0f / 255.0f), t10 = float(px[4 * (size_t(y0) * w + x1) + c]) * (1.0f / 255.0f);
        const float t01 = float(px[4 * (size_t(y1) * w + x0) + c]) * (1.0f / 255.0f), t11 = float(px[4 * (size_t(y1) * w + x1) + c]) * (1.0f / 255.0f);
        r[c] = (1.0f - a) * (1.0f - b) * t00 + a * (1.0f - b) * t10 + (1.0f - a) * b * t01 + a * b * t11;
    }
    return mk4(r[0], r[1], r[2], r[3]);
}

// tex::lookup_float4 of texture `index1` (1-based; 0 / out of range = invalid -> zero, texture_support_cuda.h:300-304)
SB_HD float4 sample_texture(const SceneDev& S, uint32_t index1, float u, float v)
{
    if (index1 == 0u || index1 > S.numTextures)
        return mk4(0.0f, 0.0f, 0.0f, 0.0f);
#if defined(__CUDA_ARCH__)
    return tex2D<float4>(cudaTextureObject_t(S.textures[index1 - 1u]), u, v);
#else
    const TexHost& t = S.texHost[index1 - 1u];
    return texture_bilinear_rgba8(t.pixels, t.width, t.height, u, v);
#endif
}

// TEX = true: the scene has textures (UsdUVTexture diffuse colour / tangent-space normal maps)
template <bool CURVES = true, bool PREVIEW = true, bool RECT_UNIFORM = false, bool HAIR = true, bool TEX = true, class Sink>
SB_HD bool shade_bounce(const FrameParams& P, const SceneDev& S, PathState& ps, const float4& ha, uint32_t hb, uint32_t depth, const uint32_t* sobolTab,
                        const float* unpackLut, Sink& sink)
{
    const uint32_t pathId = ps.pathId;
    const float3 rayO = ps.o, rayD = ps.d;
    float3 throughput = ps.throughput;
    uint32_t flags = ps.flags;
    const float lastBsdfPdf = ps.lastBsdfPdf;
    const uint32_t kind = hb >> 30;
    if (kind == 0u)
        return false; // __miss__ms: radiance += throughput * bg_color(0); path ends
    // independent loads issued together: instance record, shading record of the triangle
    // (references, not copies, for the instance / material / light records: the register-bound shade kernel then
    // re-reads fields from L1 instead of spilling them -- measured -6 %)
    const InstDev& I = S.instances[hb & 0x0fffffffu];
    TriShadeRegs tri;
    tri.a = tri.b = tri.c = tri.d = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    if (kind == 1u)
        tri = load_tri_shade(&S.triShade[f2u(ha.w)]);
    float3 Lpath = mk3(ps.L);

    if (I.type == SB_INSTANCE_LIGHT)
    {
        // __closesthit__light, OptixRender.cu:315-341 (quirks Q5, Q19)
        if (I.light < S.numLights)
        {
            const sb_light& l = S.lights[I.light];
            const float3 hitPoint = rayO + ha.x * rayD;
            const float3 ln = light_normal(l, hitPoint);
            const float3 color = mk3(l.color[0], l.color[1], l.color[2]);
            if (-dot(rayD, ln) > 0.0f)
            {
                if (depth == 0u || (flags & kFlagSpecular))
                {
                    Lpath += throughput * color * -dot(rayD, ln);
                }
                else
                {
                    const float lightPdf = light_pdf<RECT_UNIFORM>(l, hitPoint, rayO) / float(S.numLights);
                    const float w = mis_balance(lastBsdfPdf, lightPdf);
                    Lpath += throughput * color * -dot(rayD, ln) * w;
                }
                ps.L = mk4(Lpath, ps.L.w);
                sink.radiance_changed(ps);
            }
        }
        return false; // throughput = 0: the path ends
    }

    // ---- __closesthit__radiance, closest_hit.cu:456-606 --------------------------------------------
    const bool isInside = (flags & kFlagInside) != 0u;
    const Surface sf = (!CURVES || kind == 1u) ? tri_surface(I, tri, ha.y, ha.z, isInside, unpackLut) : curve_surface(S, I, f2u(ha.w), ha.y, ha.x, rayO, rayD, isInside);
    // (measured: fetching the material or computing the Sobol values earlier, to overlap them with the fetches
    // above, lengthens live ranges in this register-bound kernel and costs 3-8 %)
    const sb_material& mat = S.materials[I.material];
    float3 baseColor = mk3(mat.base_color[0], mat.base_color[1], mat.base_color[2]);
    float3 shadingNormal = sf.normal;
    if (TEX && kind == 1u && S.triUv != nullptr && (mat.diffuse_texture | mat.normal_texture) != 0u)
    {
        // state.text_coords[0] (closest_hit.cu:391-396); tangent_u / tangent_v = world tangent / cross(N, T) (:408, 485-486)
        const uint4 puv = S.triUv[f2u(ha.w)];
        const float2 uv0 = unpack_uv(puv.x), uv1 = unpack_uv(puv.y), uv2 = unpack_uv(puv.z);
        const float w0 = 1.0f - ha.y - ha.z;
        const float tu = uv0.x * w0 + uv1.x * ha.y + uv2.x * ha.z, tv = uv0.y * w0 + uv1.y * ha.y + uv2.y * ha.z;
        if (mat.diffuse_texture)
            baseColor = mk3(sample_texture(S, mat.diffuse_texture, tu, tv));
        if (mat.normal_texture)
        {
            // UsdUVTexture normal map: scale 2, bias -1; tangent space (tangent_u, tangent_v, normal)
            const float4 t = sample_texture(S, mat.normal_texture, tu, tv);
            const float3 nts = mk3(t.x * 2.0f - 1.0f, t.y * 2.0f - 1.0f, t.z * 2.0f - 1.0f);
            const float3 bitangent = cross(sf.normal, sf.tangent);
            const float3 nw = nts.x * sf.tangent + nts.y * bitangent + nts.z * sf.normal;
            if (dot(nw, nw) > 0.0f)
                shadingNormal = normalize(nw);
        }
    }
    if (P.debug == 1u)
    {
        // closest_hit.cu:504-508, after mdlcode_init (:502) has put the material's shading normal into state.normal
        ps.L = mk4((shadingNormal + mk3(1.0f)) * 0.5f, 0.0f);
        sink.radiance_changed(ps);
        return false;
    }
    uint32_t px, py, pk;
    path_pixel(P, pathId, px, py, pk);
    const uint32_t sidx = sampler_index(px, py, P.sampleBase + pk * P.sampleStride, P.sppTotal);
    // the five distinct Sobol values of this bounce (quirk Q2): xi = v[0..3], lightId = v[2],
    // lightPoint = (v[3], v[4]), russian roulette = v[4]
    const Sample5 rn = sampler_sample5(sidx, depth, sobolTab);
    const float3 k1 = -rayD;
    // hair: material constants, fibre frame and everything that depends on k1 only are shared by the sample and the NEE
    // evaluation of this bounce (hair.cuh; same arithmetic as the one-shot forms)
    const bool isHair = HAIR && mat.model == SB_MATERIAL_HAIR;
    HairCtx hc;
    BsdfSample bs;
    if (isHair)
    {
        hc = hair_prepare(mat, shadingNormal, sf.tangent, k1);
        bs.k2 = mk3(0.0f);
        bs.bsdf_over_pdf = mk3(0.0f);
        bs.pdf = 0.0f;
        bs.event = EV_ABSORB;
        if (hair_sample(hc, mk4(rn.v[0], rn.v[1], rn.v[2], rn.v[3]), bs.k2, bs.bsdf_over_pdf, bs.pdf))
            bs.event = EV_GLOSSY | (dot(sf.geomNormal, bs.k2) >= 0.0f ? EV_REFLECTION : EV_TRANSMISSION);
    }
    else
        bs = bsdf_sample<PREVIEW, false>(mat, baseColor, shadingNormal, sf.geomNormal, sf.tangent, k1, mk4(rn.v[0], rn.v[1], rn.v[2], rn.v[3]));
    if (bs.event == EV_ABSORB)
    {
        return false; // throughput = 0 (firstEventType = eAbsorb: counted by neither AOV)
    }
    const bool specularBounce = (bs.event & EV_SPECULAR) != 0;
    if (depth == 0u)
    {
        uint32_t ev = 0u;
        if (bs.event & EV_DIFFUSE)
            ev = 2u;
        if (bs.event & EV_GLOSSY)
            ev = 3u;
        flags = (flags & ~(3u << kFlagEventShift)) | (ev << kFlagEventShift);
        // AOV views (debug 2 / 3): remember the first event of this sample next to its radiance; nothing has
        // been added to L yet at depth 0 (NEE contributions arrive with the shadow ray)
        if (P.debug >= 2u)
        {
            ps.L = mk4(0.0f, 0.0f, 0.0f, u2f(ev));
            sink.radiance_changed(ps);
        }
    }
    if (bs.event & (EV_DIFFUSE | EV_GLOSSY))
    {
        // estimateDirectLighting + sampleLight, closest_hit.cu:260-324
        if (S.numLights > 0u) // numLights == 0 is undefined behaviour in the reference (quirk Q18)
        {
            uint32_t lightId = uint32_t(float(S.numLights) * rn.v[2]);
            if (lightId >= S.numLights)
                lightId = S.numLights - 1u;
            const float lightSelectionPdf = 1.0f / float(S.numLights);
            const sb_light& l = S.lights[lightId];
            const LightSample ls = sample_light<RECT_UNIFORM>(l, rn.v[3], rn.v[4], sf.position, P.rectMethod);
            const float3 Li = mk3(l.color[0], l.color[1], l.color[2]);
            if (dot(shadingNormal, ls.L) > 0.0f && -dot(ls.L, ls.normal) > 0.0 && all_nonzero(Li))
            {
                const float lightPdf = ls.pdf * lightSelectionPdf;
                const float3 radiance = Li * saturate(dot(shadingNormal, ls.L)); // visibility applied by the shadow ray
                if (isnan3(radiance) || isnanf_(lightPdf))
                {
                    ps.L = mk4(10000.0f, 0.0f, 0.0f, 0.0f); // quirk Q17
                    sink.radiance_changed(ps);
                    return false;
                }
                const bool nextEventValid = ((dot(ls.L, shadingNormal) > 0.0f) != isInside) && lightPdf != 0.0f;
                if (nextEventValid)
                {
                    BsdfEval ev;
                    if (isHair)
                    {
                        ev.diffuse = mk3(0.0f);
                        hair_evaluate(hc, ls.L, ev.glossy, ev.pdf);
                    }
                    else
                        ev = bsdf_evaluate<PREVIEW, false>(mat, baseColor, shadingNormal, sf.geomNormal, sf.tangent, k1, ls.L);
                    if (isnan3(ev.diffuse) || isnan3(ev.glossy))
                    {
                        ps.L = mk4(10000.0f, 0.0f, 0.0f, 0.0f);
                        sink.radiance_changed(ps);
                        return false;
                    }
                    if (ev.pdf > 0.0f)
                    {
                        const float3 radianceOverPdf = radiance / lightPdf;
                        const float w = mis_balance(lightPdf, ev.pdf);
                        const float3 contrib = throughput * radianceOverPdf * w * (ev.diffuse + ev.glossy);
                        if (contrib.x != 0.0f || contrib.y != 0.0f || contrib.z != 0.0f)
                        {
                            sink.shadow_ray(ps, mk4(offset_ray(sf.position, sf.geomNormal), P.shadowTmin), mk4(ls.L, ls.distToLight), contrib);
                        }
                    }
                }
            }
        }
    }
    // next path segment, closest_hit.cu:591-605
    float3 newOrigin;
    if (bs.event & EV_TRANSMISSION)
    {
        flags ^= kFlagInside;
        newOrigin = offset_ray(sf.position, -sf.geomNormal);
    }
    else
    {
        newOrigin = offset_ray(sf.position, sf.geomNormal);
    }
    flags = specularBounce ? (flags | kFlagSpecular) : (flags & ~kFlagSpecular);
    const float newPdf = specularBounce ? 1.0f : bs.pdf; // quirk Q19
    throughput *= bs.bsdf_over_pdf;
    // tail of the bounce loop, OptixRender.cu:134-153
    if (depth > 3u)
    {
        const float p = maxcomp(throughput);
        if (rn.v[4] > p)
            return false;
        throughput *= 1.0f / (p + 1e-5f);
    }
    if (dot(throughput, throughput) < 1e-5f)
        return false;
    if (depth + 1u >= P.maxDepth)
        return false;
    ps.o = newOrigin;
    ps.d = bs.k2;
    ps.lastBsdfPdf = newPdf;
    ps.throughput = throughput;
    ps.flags = flags;
    return true;
}

// Wavefront form of one bounce: path state in, shadow-ray and next-segment queue records out.
struct QueueSink
{
    const Queues& Q;
    uint32_t depth;
    SB_HD void radiance_changed(const PathState& ps) const
    {
        Q.Lacc[ps.pathId] = ps.L;
    }
    SB_HD void shadow_ray(const PathState& ps, const float4& o, const float4& d, const float3& contrib) const
    {
        const uint32_t sslot = queue_alloc(&Q.counts[count_shadow(depth)]);
        Q.shO[sslot] = o;
        Q.shD[sslot] = d;
        Q.shC[sslot] = mk4(contrib, u2f(ps.pathId));
    }
};
SB_HD void shade_one(const FrameParams& P, const SceneDev& S, const Queues& Q, uint32_t depth, int qi, uint32_t slot, const uint32_t* sobolTab,
                     const float* unpackLut)
{
    const int qo = qi ^ 1;
    const float4 ro = qsel(Q.rayO, qi)[slot], rd = qsel(Q.rayD, qi)[slot], th = qsel(Q.thr, qi)[slot];
    const float4 ha = Q.hitA[slot];
    const uint32_t hb = Q.hitB[slot];
    if ((hb >> 30) == 0u)
        return; // miss
    PathState ps;
    ps.o = mk3(ro);
    ps.d = mk3(rd);
    ps.lastBsdfPdf = rd.w;
    ps.throughput = mk3(th);
    ps.flags = f2u(th.w);
    ps.pathId = f2u(ro.w);
    ps.L = Q.Lacc[ps.pathId];
    QueueSink sink = { Q, depth };
    if (shade_bounce(P, S, ps, ha, hb, depth, sobolTab, unpackLut, sink))
    {
        const uint32_t nslot = queue_alloc(&Q.counts[count_path(depth + 1u)]);
        qsel(Q.rayO, qo)[nslot] = mk4(ps.o, u2f(ps.pathId));
        qsel(Q.rayD, qo)[nslot] = mk4(ps.d, ps.lastBsdfPdf);
        qsel(Q.thr, qo)[nslot] = mk4(ps.throughput, u2f(ps.flags));
    }
}

// closest hit of one ray: triangles first, then curves (strictly closer only).
// ha = t, u, v, bits(global triangle id | curve SegInfo index); hb = instance | kind << 30 (kind 0 = miss)
// CURVES = false compiles the curve BVH out (scenes without curves: the traversal kernels are sensitive to code size)
template <bool STATS, bool CURVES = true>
SB_HD void trace_closest(const SceneDev& S, const float3& o, const float3& d, float tmin, float4& ha, uint32_t& hb, TravStats* st)
{
    Ray ray;
    ray.o = o;
    ray.d = d;
    ray.tmin = tmin;
    ray.tmax = 1e16f;
    const RayPrep rp = prepare_ray(ray.d);
    HitRec hit;
    hit.t = 0.0f;
    hit.u = hit.v = 0.0f;
    hit.prim = hit.inst = hit.kind = 0u;
    hit.gid = 0xffffffffu;
    if (S.numTriNodes)
        traverse_bvh<1, false, STATS>(S.triNodes, S.tris, kRayMaskPrimary, ray, rp, hit, st);
    if (CURVES && S.numSegNodes)
    {
        if (traverse_bvh<2, false, STATS>(S.segNodes, S.segs, kRayMaskPrimary, ray, rp, hit, st))
        {
            const SegInfo si = S.segInfo[hit.prim];
            hit.inst = si.inst;
            hit.u = span_to_segment_u(si.span, hit.u);
        }
    }
    ha = mk4(hit.t, hit.u, hit.v, u2f(hit.kind == 1u ? hit.gid : hit.prim));
    hb = hit.inst | (hit.kind << 30);
}

// any hit along a shadow ray (so = origin.xyz, tmin; sd = dir.xyz, tmax)
template <bool STATS, bool CURVES = true>
SB_HD bool trace_occluded(const SceneDev& S, const float4& so, const float4& sd, TravStats* st)
{
    Ray ray;
    ray.o = mk3(so);
    ray.tmin = so.w;
    ray.d = mk3(sd);
    ray.tmax = sd.w;
    const RayPrep rp = prepare_ray(ray.d);
    HitRec hit;
    hit.gid = 0xffffffffu;
    hit.kind = 0u;
    bool occluded = false;
    if (S.numTriNodes)
        occluded = traverse_bvh<1, true, STATS>(S.triNodes, S.tris, kRayMaskShadow, ray, rp, hit, st);
    if (CURVES && !occluded && S.numSegNodes)
        occluded = traverse_bvh<2, true, STATS>(S.segNodes, S.segs, kRayMaskShadow, ray, rp, hit, st);
    return occluded;
}

template <bool STATS, bool CURVES = true>
SB_HD void extend_one(const FrameParams& P, const SceneDev& S, const Queues& Q, int qi, uint32_t slot, TravStats* st)
{
    const float4 ro = qsel(Q.rayO, qi)[slot], rd = qsel(Q.rayD, qi)[slot];
    float4 ha;
    uint32_t hb;
    trace_closest<STATS, CURVES>(S, mk3(ro), mk3(rd), P.materialTmin, ha, hb, st);
    Q.hitA[slot] = ha;
    Q.hitB[slot] = hb;
}

template <bool STATS, bool CURVES = true>
SB_HD void shadow_one(const SceneDev& S, const Queues& Q, uint32_t j, TravStats* st)
{
    const float4 so = Q.shO[j], sd = Q.shD[j], sc = Q.shC[j];
    if (!trace_occluded<STATS, CURVES>(S, so, sd, st))
    {
        const uint32_t pathId = f2u(sc.w);
        const float4 L = Q.Lacc[pathId];
        Q.Lacc[pathId] = mk4(L.x + sc.x, L.y + sc.y, L.z + sc.z, L.w);
    }
}

// ---- fused form: one whole bounce (closest hit, shading, shadow ray) with the path state in registers ------
// Used by the single-kernel path tracer that small scenes run (kernels.cu: k_path_fused): when the whole BVH is a
// handful of nodes, the queue traffic of the wavefront form costs more than the divergence it removes.  Same
// functions, same order of floating-point operations per path as the wavefront form: identical images.
struct RegisterSink
{
    bool shadow;
    float4 o, d;
    float3 contrib;
    SB_HD void radiance_changed(const PathState&) const
    {
    }
    SB_HD void shadow_ray(const PathState&, const float4& so, const float4& sd, const float3& c)
    {
        shadow = true;
        o = so;
        d = sd;
        contrib = c;
    }
};

// returns true when the path continues at depth + 1
template <bool STATS>
SB_HD bool path_bounce(const FrameParams& P, const SceneDev& S, PathState& ps, uint32_t depth, const uint32_t* sobolTab, const float* unpackLut,
                       uint32_t& shadowRays, TravStats* stExtend, TravStats* stShadow)
{
    float4 ha;
    uint32_t hb;
    trace_closest<STATS>(S, ps.o, ps.d, P.materialTmin, ha, hb, stExtend);
    RegisterSink sink;
    sink.shadow = false;
    const bool next = shade_bounce(P, S, ps, ha, hb, depth, sobolTab, unpackLut, sink);
    if (sink.shadow)
    {
        ++shadowRays;
        if (!trace_occluded<STATS>(S, sink.o, sink.d, stShadow))
            ps.L = mk4(ps.L.x + sink.contrib.x, ps.L.y + sink.contrib.y, ps.L.z + sink.contrib.z, ps.L.w);
    }
    return next;
}

// tonemap / inverseTonemap, postprocessing/Utils.h:5-15
SB_HD float3 tonemap3(float3 c, const float3& e)
{
    c = c * e;
    return c / (c + mk3(1.0f));
}
SB_HD float3 inverse_tonemap3(const float3& c, const float3& e)
{
    return c / (e - c * e);
}

// Fold the finished samples of one (padded) pixel into the accumulation buffer.
//  mode 0: S += sum_k T(L_k)                      (spp == 1 per launch: the reference's running mean of T)
//  mode 1: reference lerp for a launch of `chunk` samples (quirk Q1: weight 1/(subframe+1))
//  mode 2: no accumulation: `direct` receives the linear mean of the launch
//  Modes 1 / 2 form the linear mean of the WHOLE launch (render/pt/spp samples).  A launch larger than one wavefront
//  batch arrives in several calls: batchFlags bit 0 = first batch of the launch, bit 1 = last; launchSamples = samples
//  of the whole launch.  Between batches the running linear sum waits in `direct` (beauty) and in scrD / scrS (the
//  AOV sums, w = matching samples so far); the additions happen in the same order as in a single batch, so the
//  result does not depend on the batch size bit for bit.
constexpr uint32_t kBatchFirst = 1u, kBatchLast = 2u;
SB_HD void accumulate_pixel(const FrameParams& P, const Queues& Q, float4* S, float4* direct, float4* aovD, float4* aovS, uint32_t mode,
                            uint32_t subframe, uint32_t p, uint32_t launchSamples = 0u, uint32_t batchFlags = kBatchFirst | kBatchLast,
                            float4* scrD = nullptr, float4* scrS = nullptr)
{
    if (launchSamples == 0u)
        launchSamples = P.chunk;
    const bool firstBatch = (batchFlags & kBatchFirst) != 0u, lastBatch = (batchFlags & kBatchLast) != 0u;
    const uint32_t tile = p >> 5, within = p & 31u;
    const uint32_t x = (tile % P.tilesX) * 8u + (within & 7u), y = (tile / P.tilesX) * 4u + (within >> 3);
    if (x >= P.width || y >= P.height)
        return;
    const uint32_t lin = y * P.width + x;
    const float3 e = mk3(P.exposure[0], P.exposure[1], P.exposure[2]);
    if (P.debug >= 2u)
    {
        // diffuse / specular AOVs, OptixRender.cu:157-221: samples whose FIRST bsdf event was diffuse / glossy,
        // accumulated like the beauty buffer but with their own (uint16, quirk Q20) sample counters.  Stored as
        // (count * A) in xyz and the count in w, A being the tone-mapped running value.  The beauty buffer is
        // not updated in these views (the reference returns early, OptixRender.cu:212-221).
        for (int which = 0; which < 2; ++which)
        {
            float4* buf = which == 0 ? aovD : aovS;
            const uint32_t wantEv = which == 0 ? 2u : 3u;
            float4 acc = buf[lin];
            uint32_t cnt = uint32_t(acc.w);
            if (mode == 0u)
            {
                for (uint32_t k = 0; k < P.chunk; ++k)
                {
                    const float4 L = Q.Lacc[k * P.nPixPadded + p];
                    if (f2u(L.w) != wantEv)
                        continue;
                    const uint32_t prev = (subframe + k > 0u) ? cnt : 0u;
                    const float3 t = tonemap3(mk3(L), e);
                    const float3 sum = prev > 0u ? mk3(acc) + t : t;
                    cnt = (prev + 1u) & 0xffffu;
                    acc = mk4(sum, float(cnt));
                }
            }
            else
            {
                float4* scr = which == 0 ? scrD : scrS;
                float3 mean = mk3(0.0f);
                uint32_t n = 0;
                if (!firstBatch)
                {
                    const float4 part = scr[lin];
                    mean = mk3(part);
                    n = uint32_t(part.w);
                }
                for (uint32_t k = 0; k < P.chunk; ++k)
                {
                    const float4 L = Q.Lacc[k * P.nPixPadded + p];
                    if (f2u(L.w) == wantEv)
                    {
                        mean += mk3(L);
                        ++n;
                    }
                }
                if (!lastBatch)
                {
                    scr[lin] = mk4(mean, float(n));
                    continue;
                }
                if (n > 0u)
                {
                    mean = mean / float(n);
                    const uint32_t prev = subframe > 0u ? cnt : 0u;
                    float3 A = tonemap3(mean, e);
                    if (prev > 0u)
                        A = lerp(mk3(acc) / float(prev), A, 1.0f / float(prev + 1u));
                    cnt = (prev + n) & 0xffffu;
                    acc = mk4(A * float(cnt), float(cnt));
                }
            }
            buf[lin] = acc;
        }
        return;
    }
    if (mode == 0u)
    {
        float3 s = mk3(S[lin]);
        for (uint32_t k = 0; k < P.chunk; ++k)
            s += tonemap3(mk3(Q.Lacc[k * P.nPixPadded + p]), e);
        S[lin] = mk4(s, 0.0f);
        return;
    }
    float3 result = firstBatch ? mk3(0.0f) : mk3(direct[lin]);
    for (uint32_t k = 0; k < P.chunk; ++k)
        result += mk3(Q.Lacc[k * P.nPixPadded + p]);
    if (!lastBatch)
    {
        direct[lin] = mk4(result, 0.0f);
        return;
    }
    result = result / float(launchSamples);
    if (mode == 2u)
    {
        direct[lin] = mk4(result, 1.0f);
        return;
    }
    // S holds n * A with A the tone-mapped running value; the reference lerps A towards T(result)
    float3 A = tonemap3(result, e);
    if (subframe > 0u)
    {
        const float3 prev = mk3(S[lin]) / float(subframe);
        const float a = 1.0f / float(subframe + 1u);
        A = lerp(prev, A, a);
    }
    S[lin] = mk4(A * float(subframe + launchSamples), 0.0f);
}

// the optional post-process of OptixRender.cpp:1045-1049: tone curve (Tonemappers.cu:17-109), then gamma
SB_HD float4 postprocess_pixel(float3 c, const float3& e, uint32_t tonemapper, float gamma);

// image = T^-1(S / n), then the post-process.  n == 0xffffffff: the count is per pixel, in s.w (AOV buffers)
SB_HD float4 resolve_pixel(const float4& s, uint32_t n, const float3& e, uint32_t tonemapper, float gamma)
{
    if (n == 0xffffffffu)
        n = uint32_t(s.w);
    float3 c = mk3(0.0f);
    if (n > 0u)
        c = inverse_tonemap3(mk3(s) / float(n), e);
    return postprocess_pixel(c, e, tonemapper, gamma);
}

SB_HD float4 postprocess_pixel(float3 c, const float3& e, uint32_t tonemapper, float gamma)
{
    if (tonemapper == 1u)
    {
        const float3 r = c * e;
        const float lum = r.x * 0.299f + r.y * 0.587f + r.z * 0.114f;
        c = r / (lum + 1);
    }
    else if (tonemapper == 2u)
    {
        const float3 r = c * e;
        float3 v = mk3(0.59719f * r.x + 0.35458f * r.y + 0.04823f * r.z, 0.07600f * r.x + 0.90834f * r.y + 0.01566f * r.z,
                       0.02840f * r.x + 0.13383f * r.y + 0.83777f * r.z);
        const float3 a = v * (v + mk3(0.0245786f)) - mk3(0.000090537f);
        const float3 b = v * (0.983729f * v + mk3(0.4329510f)) + mk3(0.238081f);
        v = a / b;
        c = mk3(1.60475f * v.x + -0.53108f * v.y + -0.07367f * v.z, -0.10208f * v.x + 1.10813f * v.y + -0.00605f * v.z,
                -0.00327f * v.x + -0.07276f * v.y + 1.07602f * v.z);
        c = mk3(saturate(c.x), saturate(c.y), saturate(c.z));
    }
    else if (tonemapper == 3u)
    {
        const float3 x = c * e;
        const float A = 2.51f, B = 0.03f, C = 2.43f, D = 0.59f, E = 0.14f;
        const float3 r = (x * (A * x + mk3(B))) / (x * (C * x + mk3(D)) + mk3(E));
        c = mk3(saturate(r.x), saturate(r.y), saturate(r.z));
    }
    if (gamma > 0.0f)
    {
        const float ig = 1.0f / gamma;
        c = mk3(powf(c.x, ig), powf(c.y, ig), powf(c.z, ig));
    }
    return mk4(c, 1.0f);
}

} // namespace sb
