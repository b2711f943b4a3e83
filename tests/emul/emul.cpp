// TEST HARNESS ONLY -- serial host instantiation of the backend's SB_HD device logic.
//
// Compiled by g++ with -DSB_HOST_EMUL (exec.h then selects the serial ExecHost policy): the SAME
// bodies that nvcc compiles into the sm_100a kernels (BVH builder steps, CWBVH traversal, raygen /
// shade / shadow / accumulate) run here as plain loops, so the CPU test-suite can compare them with
// the oracle before any GPU time is spent.  Mirrors launch_wavefront_batch()/render_samples() of
// kernels.cu / sb_api.cu.  Never linked into libstrelka_b200.so; nothing in the product calls it.
#define SB_HOST_EMUL 1
#include <cstdint>
namespace sb
{
uint32_t h_sobol[5][32];
}
#include "../../strelka_b200/csrc/wavefront.cuh"
#include "../../strelka_b200/csrc/scene_prep.h"
#include "../../strelka_b200/csrc/host_math.h"

#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

using namespace sb;

struct emul_scene
{
    SceneDev S;
    std::vector<void*> owned;
};

template <class T>
static T* dup(emul_scene* e, const T* src, size_t n)
{
    T* p = static_cast<T*>(std::calloc(n ? n : 1, sizeof(T)));
    if (n)
        std::memcpy(p, src, n * sizeof(T));
    e->owned.push_back(p);
    return p;
}

extern "C" {

emul_scene* emul_scene_create(const sb_scene_view* v, uint32_t curveSplit, char* err, int errLen)
{
    sobol_generate(h_sobol);
    emul_scene* e = new emul_scene();
    try
    {
        ScenePrep prep;
        prepare_scene(v, prep);
        SceneDev& s = e->S;
        s.vertices = dup(e, v->vertices, v->num_vertices);
        s.indices = dup(e, v->indices, v->num_indices);
        s.meshes = dup(e, v->meshes, v->num_meshes);
        s.curves = dup(e, v->curves, v->num_curves);
        s.curvePoints = dup(e, v->curve_points, v->num_curve_points * 3);
        s.curveRadii = dup(e, v->curve_widths, v->num_curve_widths);
        s.curveVertexCounts = dup(e, v->curve_vertex_counts, v->num_curve_vertex_counts);
        s.lights = dup(e, v->lights, v->num_lights);
        if (v->num_materials)
        {
            s.materials = dup(e, v->materials, v->num_materials);
        }
        else
        {
            sb_material def;
            std::memset(&def, 0, sizeof(def));
            def.base_color[0] = def.base_color[1] = def.base_color[2] = 1.0f;
            s.materials = dup(e, &def, 1);
        }
        s.instances = dup(e, prep.inst.data(), prep.inst.size());
        s.numInstances = v->num_instances;
        s.numLights = v->num_lights;
        s.numMaterials = prep.numMaterials;
        s.numMeshes = v->num_meshes;
        s.numCurves = v->num_curves;
        s.numCurvePoints = v->num_curve_points;
        s.numCurveRadii = v->num_curve_widths;
        // textures: the host emulation samples private copies with the documented filter (wavefront.cuh)
        if (v->num_textures && v->textures)
        {
            TexHost* th = static_cast<TexHost*>(std::calloc(v->num_textures, sizeof(TexHost)));
            e->owned.push_back(th);
            bool used = false;
            for (uint32_t i = 0; i < v->num_textures; ++i)
            {
                th[i].width = v->textures[i].width;
                th[i].height = v->textures[i].height;
                th[i].pixels = dup(e, v->textures[i].pixels, size_t(th[i].width) * th[i].height * 4);
            }
            for (uint32_t i = 0; i < v->num_materials; ++i)
                used = used || v->materials[i].diffuse_texture || v->materials[i].normal_texture;
            s.texHost = th;
            s.numTextures = v->num_textures;
            if (used && prep.numTris)
            {
                s.triUv = static_cast<uint4*>(std::calloc(prep.numTris, sizeof(uint4)));
                e->owned.push_back(s.triUv);
            }
        }
        uint32_t* triFirst = dup(e, prep.triFirst.data(), prep.triFirst.size());
        SegInfo* segInfo = dup(e, prep.segInfo.data(), prep.segInfo.size());
        ExecHost ex;
        build_scene_bvhs(ex, s, triFirst, uint32_t(prep.numTris), nullptr, uint32_t(prep.segInfo.size()), segInfo, curveSplit);
    }
    catch (const std::exception& ex)
    {
        if (err && errLen > 0)
        {
            std::strncpy(err, ex.what(), size_t(errLen) - 1);
            err[errLen - 1] = 0;
        }
        delete e;
        return nullptr;
    }
    return e;
}

void emul_scene_destroy(emul_scene* e)
{
    if (!e)
        return;
    for (void* p : e->owned)
        std::free(p);
    std::free(e->S.tris);
    std::free(e->S.triShade);
    std::free(e->S.segs);
    std::free(e->S.segInfo);
    std::free(e->S.triNodes);
    std::free(e->S.segNodes);
    delete e;
}

void emul_scene_info(const emul_scene* e, uint64_t* out /* [4] tris, segs, triNodes, segNodes */)
{
    out[0] = e->S.numTris;
    out[1] = e->S.numSegs;
    out[2] = e->S.numTriNodes;
    out[3] = e->S.numSegNodes;
}

// Structure of the triangle BVH: walks every wide node from the root.
// out[0] nodes reached, [1] inner children, [2] leaf slots, [3] primitive references, [4] primitives referenced
// more than once or never (must be 0), [5] deepest level, [6] empty child slots
void emul_bvh_stats(const emul_scene* e, uint64_t* out)
{
    for (int i = 0; i < 7; ++i)
        out[i] = 0;
    const SceneDev& S = e->S;
    if (!S.numTriNodes)
        return;
    std::vector<uint32_t> seen(S.numTris, 0u);
    std::vector<std::pair<uint32_t, uint32_t>> stack; // node, level
    stack.emplace_back(0u, 1u);
    while (!stack.empty())
    {
        const auto [ni, level] = stack.back();
        stack.pop_back();
        const WideNode& n = S.triNodes[ni];
        out[0]++;
        out[5] = std::max<uint64_t>(out[5], level);
        const uint32_t imask = n.n0.w >> 24, childBase = n.n1.x, primBase = n.n1.y;
        uint32_t innerSeen = 0;
        for (int j = 0; j < 8; ++j)
        {
#if SB_FIXED_BITS
            // valid word: slot j owns bits 3j..3j+2 (leaf, unary primitive count) or bit 24+j (inner node)
            const uint32_t valid = n.n1.z;
            const bool inner = ((valid >> (24 + j)) & 1u) != 0u;
            const uint32_t childBits = (valid >> (3 * j)) & 7u;
            const uint32_t bitIndex = inner ? 24u : uint32_t(__builtin_popcount(valid & ((1u << (3 * j)) - 1u)));
            if (!inner && childBits == 0u)
            {
                out[6]++;
                continue;
            }
            if (inner && childBits != 0u)
                out[4] += 1u << 21; // a slot cannot be both
#else
            const uint32_t meta = ((j < 4 ? n.n1.z : n.n1.w) >> (8 * (j & 3))) & 0xffu;
            if (meta == 0u)
            {
                out[6]++;
                continue;
            }
            const uint32_t bitIndex = meta & 31u, childBits = meta >> 5;
#endif
            if (bitIndex >= 24u)
            {
                out[1]++;
                // inner children are stored in slot order: rank among the inner slots below this one
                const uint32_t rel = uint32_t(__builtin_popcount(imask & ((1u << j) - 1u)));
                stack.emplace_back(childBase + rel, level + 1u);
                innerSeen |= 1u << j;
            }
            else
            {
                out[2]++;
                for (uint32_t b = 0; b < 3; ++b)
                    if (childBits & (1u << b))
                    {
                        out[3]++;
                        const uint32_t gid = f2u(S.tris[primBase + bitIndex + b].e2.w);
                        seen[gid]++;
                    }
            }
        }
        if (innerSeen != imask)
            out[4] += 1u << 20; // inner mask and meta bytes disagree
    }
    for (uint32_t c : seen)
        if (c != 1u)
            out[4]++;
}

// same contract as sb_test_trace; stats (optional) [nodes, tris, segs, overflow] summed over rays
void emul_trace(const emul_scene* e, uint32_t n, const float* rays, uint32_t mode, sb_hit* hits, uint64_t* stats)
{
    const SceneDev& S = e->S;
    for (uint32_t i = 0; i < n; ++i)
    {
        const float* r = rays + 8 * size_t(i);
        Ray ray;
        ray.o = mk3(r[0], r[1], r[2]);
        ray.tmin = r[3];
        ray.d = mk3(r[4], r[5], r[6]);
        ray.tmax = r[7];
        const RayPrep rp = prepare_ray(ray.d);
        HitRec hit;
        hit.t = hit.u = hit.v = 0.0f;
        hit.prim = hit.inst = hit.kind = 0u;
        hit.gid = 0xffffffffu;
        TravStats st = { 0, 0, 0, 0 };
        sb_hit h = { 0, 0, 0, 0, 0, 0 };
        if (mode == 0)
        {
            if (S.numTriNodes)
                traverse_bvh<1, false, true>(S.triNodes, S.tris, kRayMaskPrimary, ray, rp, hit, &st);
            if (S.numSegNodes)
            {
                if (traverse_bvh<2, false, true>(S.segNodes, S.segs, kRayMaskPrimary, ray, rp, hit, &st))
                {
                    const SegInfo si = S.segInfo[hit.prim];
                    hit.inst = si.inst;
                    hit.prim = si.prim;
                    hit.u = span_to_segment_u(si.span, hit.u);
                }
            }
            h.t = hit.t;
            h.u = hit.u;
            h.v = hit.v;
            h.prim = hit.prim;
            h.instance = hit.inst;
            h.kind = hit.kind;
        }
        else
        {
            bool occ = false;
            if (S.numTriNodes)
                occ = traverse_bvh<1, true, true>(S.triNodes, S.tris, kRayMaskShadow, ray, rp, hit, &st);
            if (!occ && S.numSegNodes)
                occ = traverse_bvh<2, true, true>(S.segNodes, S.segs, kRayMaskShadow, ray, rp, hit, &st);
            h.kind = occ ? 1u : 0u;
        }
        hits[i] = h;
        if (stats)
        {
            stats[0] += st.nodes;
            stats[1] += st.tris;
            stats[2] += st.segs;
            stats[3] += st.overflow;
        }
    }
}

// Serial twin of render_samples()+write_output() for the spp == 1 accumulation mode: renders local
// samples [subframe, subframe + samples) into S (float4 per pixel, caller-zeroed for subframe 0) and
// resolves `image`.  counters: [paths, radiance rays, shadow rays].
void emul_render(const emul_scene* e, const sb_settings* st, const float* view, float fovY, uint32_t width, uint32_t height,
                 uint32_t subframe, uint32_t samples, uint32_t chunkMax, float* Sbuf, float* image, uint64_t* counters, uint32_t fused)
{
    const SceneDev& S = e->S;
    FrameParams P;
    std::memset(&P, 0, sizeof(P));
    P.width = width;
    P.height = height;
    P.tilesX = (width + 7) / 8;
    P.nPixPadded = P.tilesX * ((height + 3) / 4) * 32u;
    P.maxDepth = st->depth;
    P.sppTotal = st->spp_total;
    P.rectMethod = st->rect_light_sampling_method;
    P.debug = (st->debug >= 1 && st->debug <= 3) ? st->debug : 0u;
    P.shadowTmin = st->shadow_ray_tmin;
    P.materialTmin = st->material_ray_tmin;
    clip_to_view_from_fov(fovY, float(width) / float(height), P.clipToView);
    view_to_world_from_view(view, P.viewToWorld);
    compute_exposure(*st, P.exposure);
    P.sampleStride = st->sample_stride ? st->sample_stride : 1u;
    P.numLights = S.numLights;
    if (chunkMax == 0)
        chunkMax = 1;
    const size_t np = size_t(P.nPixPadded) * chunkMax;
    std::vector<float4> rayO[2], rayD[2], thr[2], hitA(np), Lacc(np), shO(np), shD(np), shC(np);
    std::vector<uint32_t> hitB(np), counts(kNumCounts);
    for (int i = 0; i < 2; ++i)
    {
        rayO[i].resize(np);
        rayD[i].resize(np);
        thr[i].resize(np);
    }
    StatCounters stats;
    std::memset(&stats, 0, sizeof(stats));
    Queues Q;
    for (int i = 0; i < 2; ++i)
    {
        Q.rayO[i] = rayO[i].data();
        Q.rayD[i] = rayD[i].data();
        Q.thr[i] = thr[i].data();
    }
    Q.hitA = hitA.data();
    Q.hitB = hitB.data();
    Q.Lacc = Lacc.data();
    Q.shO = shO.data();
    Q.shD = shD.data();
    Q.shC = shC.data();
    Q.counts = counts.data();
    Q.stats = &stats;
    std::vector<uint32_t> sobolTab(kSobolTabWords);
    sobol_build_tables(h_sobol, sobolTab.data());
    Q.sobolTab = sobolTab.data();
    std::vector<float> unpackLut(kUnpackLutSize);
    for (uint32_t i = 0; i < kUnpackLutSize; ++i)
        unpackLut[i] = unpack_component(i);
    float4* Sacc = reinterpret_cast<float4*>(Sbuf);
    std::vector<float4> direct(size_t(width) * height), aovD(size_t(width) * height, mk4(0, 0, 0, 0)), aovS(size_t(width) * height, mk4(0, 0, 0, 0));
    const bool debugNormals = P.debug == 1u;
    const uint32_t mode = debugNormals ? 2u : 0u;
    uint32_t done = 0;
    while (done < samples)
    {
        const uint32_t chunk = std::min(samples - done, chunkMax);
        P.chunk = chunk;
        P.sampleBase = st->sample_offset + (subframe + done) * P.sampleStride;
        std::fill(counts.begin(), counts.end(), 0u);
        counters[0] += uint64_t(width) * height * chunk;
        if (fused)
        {
            // the per-lane body of k_path_fused: whole paths, state in "registers"
            for (uint32_t i = 0; i < P.nPixPadded * chunk; ++i)
            {
                PathState ps;
                if (!raygen_state(P, i, ps))
                {
                    Q.Lacc[i] = mk4(0.0f, 0.0f, 0.0f, 0.0f);
                    continue;
                }
                TravStats te = { 0, 0, 0, 0 }, tsh = { 0, 0, 0, 0 };
                uint32_t depth = 0, nShadow = 0;
                for (;;)
                {
                    ++counters[1];
                    const bool next = path_bounce<false>(P, S, ps, depth, Q.sobolTab, unpackLut.data(), nShadow, &te, &tsh);
                    if (!next)
                        break;
                    ++depth;
                }
                counters[2] += nShadow;
                Q.Lacc[i] = ps.L;
            }
        }
        if (!fused)
            for (uint32_t i = 0; i < P.nPixPadded * chunk; ++i)
                raygen_one(P, Q, i);
        for (uint32_t depth = 0; depth < (fused ? 0u : P.maxDepth); ++depth)
        {
            TravStats ts = { 0, 0, 0, 0 };
            const uint32_t n = counts[count_path(depth)];
            counters[1] += n;
            for (uint32_t i = 0; i < n; ++i)
                extend_one<false>(P, S, Q, int(depth & 1u), i, &ts);
            for (uint32_t i = 0; i < n; ++i)
                shade_one(P, S, Q, depth, int(depth & 1u), i, (depth & 1u) ? Q.sobolTab : nullptr, (depth & 1u) ? unpackLut.data() : nullptr); // both code paths, same bits
            if (debugNormals)
                break;
            const uint32_t ns = counts[count_shadow(depth)];
            counters[2] += ns;
            for (uint32_t i = 0; i < ns; ++i)
                shadow_one<false>(S, Q, i, &ts);
        }
        for (uint32_t p = 0; p < P.nPixPadded; ++p)
            accumulate_pixel(P, Q, Sacc, direct.data(), aovD.data(), aovS.data(), mode, subframe + done, p);
        done += chunk;
    }
    const float3 ex = mk3(P.exposure[0], P.exposure[1], P.exposure[2]);
    float4* img = reinterpret_cast<float4*>(image);
    for (size_t i = 0; i < size_t(width) * height; ++i)
    {
        if (debugNormals)
            img[i] = direct[i];
        else if (P.debug >= 2u)
            img[i] = resolve_pixel(P.debug == 2u ? aovD[i] : aovS[i], 0xffffffffu, ex, st->tonemapper_type, st->gamma);
        else
            img[i] = resolve_pixel(Sacc[i], subframe + samples, ex, st->tonemapper_type, st->gamma);
    }
}

void emul_sampler(uint32_t n, const uint32_t* x, const uint32_t* y, const uint32_t* sample, const uint32_t* maxs, const uint32_t* depth,
                  const uint32_t* dim, float* out)
{
    sobol_generate(h_sobol);
    for (uint32_t i = 0; i < n; ++i)
    {
        const uint32_t sidx = sampler_index(x[i], y[i], sample[i], maxs[i]);
        const float a = sampler_rnd(sidx, depth[i], dim[i]);
        const Sample5 s5 = sampler_sample5(sidx, depth[i]);
        const float b = s5.v[(dim[i] + depth[i] * 10u) % 5u];
        out[i] = (a == b) ? a : -1.0f;
    }
}

void emul_light_sample(uint32_t n, const sb_light* lights, const float* hp, const float* u, uint32_t method, float* out)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const LightSample s = sample_light(lights[i], u[2 * i], u[2 * i + 1], mk3(hp[3 * i], hp[3 * i + 1], hp[3 * i + 2]), method);
        float* o = out + 12 * size_t(i);
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

// out: hit, t, u
void emul_curve_intersect(const float* q, const float* ray, float* out)
{
    float4 cp[4];
    for (int k = 0; k < 4; ++k)
        cp[k] = mk4(q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);
    float t = 0.0f, u = 0.0f;
    const CurveSpan sp = curve_span(cp, 0, 1); // the whole segment as one span
    const bool hit = intersect_round_cubic(sp.c, mk3(ray[0], ray[1], ray[2]), mk3(ray[3], ray[4], ray[5]), ray[6], ray[7], t, u);
    out[0] = hit ? 1.0f : 0.0f;
    out[1] = t;
    out[2] = u;
}

// the device BSDF code on the host: same layout as sb_test_bsdf (19 floats in, 15 out per item)
void emul_bsdf_batch(const sb_material* m, uint32_t n, const float* in, float* out)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        const float* a = in + 19 * size_t(i);
        float* o = out + 15 * size_t(i);
        const float3 N = mk3(a[0], a[1], a[2]), NG = mk3(a[3], a[4], a[5]), T = mk3(a[6], a[7], a[8]), K1 = mk3(a[9], a[10], a[11]);
        const float3 base = mk3(m->base_color[0], m->base_color[1], m->base_color[2]);
        const BsdfSample s = bsdf_sample<true, true>(*m, base, N, NG, T, K1, mk4(a[12], a[13], a[14], a[15]));
        o[0] = s.k2.x;
        o[1] = s.k2.y;
        o[2] = s.k2.z;
        o[3] = s.bsdf_over_pdf.x;
        o[4] = s.bsdf_over_pdf.y;
        o[5] = s.bsdf_over_pdf.z;
        o[6] = s.pdf;
        o[7] = float(s.event);
        const BsdfEval e = bsdf_evaluate<true, true>(*m, base, N, NG, T, K1, mk3(a[16], a[17], a[18]));
        o[8] = e.diffuse.x;
        o[9] = e.diffuse.y;
        o[10] = e.diffuse.z;
        o[11] = e.glossy.x;
        o[12] = e.glossy.y;
        o[13] = e.glossy.z;
        o[14] = e.pdf;
    }
}

// The octant permutation of the inner hit byte (traverse.cuh): the device reads the constexpr table, the host path of
// permute_inner_hits computes it bit by bit.  Returns the number of (octant, byte) pairs on which they differ, plus
// the table itself (2048 bytes) for the caller to check against its own definition.
int emul_octant_perm_table(unsigned char* table)
{
    int bad = 0;
#if SB_FIXED_BITS
    const OctantPermLut lut;
    for (uint32_t o = 0; o < 8u; ++o)
        for (uint32_t x = 0; x < 256u; ++x)
        {
            table[o * 256u + x] = lut.v[o * 256u + x];
            if (uint32_t(lut.v[o * 256u + x]) != permute_inner_hits(x, o << 8))
                ++bad;
        }
#endif
    return bad;
}

void emul_camera(const float* view, float fovY, float aspect, float* clipToView, float* viewToWorld)
{
    clip_to_view_from_fov(fovY, aspect, clipToView);
    view_to_world_from_view(view, viewToWorld);
}

} // extern "C"
