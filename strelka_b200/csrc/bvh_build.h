// On-device construction of the world-space primitive lists and their 8-wide compressed BVHs.
//
// Pipeline (all steps are data-parallel kernels; the host only sequences them and reads back one
// counter per PLOC iteration / collapse level):
//   flatten  : (instance, primitive) -> world-space TriRec / SegRec + AABB        [replaces the IAS]
//   morton   : 63-bit Morton code of the AABB centre, radix sort (cub)
//   PLOC     : parallel locally-ordered clustering -> binary BVH (Meister & Bittner 2018)
//   collapse : level-synchronous greedy BVH2 -> BVH8 conversion + quantisation (CWBVH)
//   reorder  : primitive records permuted into leaf order
// Replaces optixAccelBuild + compaction (OptixRender.cpp:300-308, 366-376, 487-492).  Deterministic:
// every allocation comes from a prefix sum, never from an atomic counter.
//
// The orchestration is written against the Exec policy of exec.h so that the CPU test-suite can run
// the identical logic serially (tests/emul); the product always instantiates ExecCuda.
#pragma once
#include "exec.h"
#include "traverse.cuh"
#include "../../include/sb/sb_api.h"
#include <algorithm>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace sb
{

struct InstDev // 128 B
{
    Affine o2w; // HitGroupData.object_to_world (OptixRender.cpp:792-793)
    Affine w2o; // HitGroupData.world_to_object = inverse (OptixRender.cpp:794-796)
    uint32_t type, geom, material, light;
    uint32_t mask, firstPrim, numPrims, pad; // firstPrim: offset of this instance in the world prim list
    float scale; // uniform scale applied to curve radii
    uint32_t pad2[3];
};

// Per-triangle shading record (64 B, one per global triangle id).
struct alignas(16) TriShade
{
    float pos[9]; // object-space positions of the three corners
    uint32_t normal[3]; // 10-10-10 packed, as in sb_vertex
    uint32_t tangent[3];
    uint32_t material; // of the owning instance
};
static_assert(sizeof(TriShade) == 64, "TriShade is fetched as four 16-byte loads");

// one texture as the host emulation sees it (the product samples cudaTextureObject_t, never this)
struct TexHost
{
    const uint8_t* pixels; // RGBA8
    uint32_t width, height;
};

struct SceneDev
{
    // uploaded scene arrays (same layouts as the host structs)
    sb_vertex* vertices = nullptr;
    uint32_t* indices = nullptr;
    sb_mesh* meshes = nullptr;
    sb_curve* curves = nullptr;
    float* curvePoints = nullptr; // 3 floats per point
    float* curveRadii = nullptr;
    uint32_t* curveVertexCounts = nullptr;
    sb_light* lights = nullptr;
    sb_material* materials = nullptr;
    InstDev* instances = nullptr;
    uint32_t numInstances = 0, numLights = 0, numMaterials = 0, numMeshes = 0, numCurves = 0;
    uint64_t numCurvePoints = 0, numCurveRadii = 0;
    // world-space geometry + BVHs
    TriRec* tris = nullptr;
    TriShade* triShade = nullptr; // per GLOBAL triangle id: everything fillTriangleGeomData reads, in one 64-byte record
    SegRec* segs = nullptr;
    SegInfo* segInfo = nullptr; // per SegRec (leaf order)
    WideNode* triNodes = nullptr;
    WideNode* segNodes = nullptr;
    uint32_t numTris = 0, numSegs = 0, numTriNodes = 0, numSegNodes = 0;
    uint32_t triDepth = 0, segDepth = 0; // levels of the two wide BVHs (<= kStackSize, checked by the builder)
    bool onlyRectLights = false; // every light is a rect light (with rectLightSamplingMethod 0 the shade kernel keeps only that sampler)
    bool anyPreviewMaterial = true; // false: every material is a diffuse one (the shade kernel drops the UsdPreviewSurface code)
    bool anyHairMaterial = false; // true: some material is SB_MATERIAL_HAIR (the shade kernel keeps the fibre BSDF)
    // textures (UsdUVTexture inputs): CUDA texture objects as the reference creates them (OptixRender.cpp:1191-1268);
    // triUv = the packed st of the three corners per GLOBAL triangle id, only built for scenes that have a texture
    const unsigned long long* textures = nullptr; // cudaTextureObject_t[numTextures] (device array)
    uint32_t numTextures = 0;
    uint4* triUv = nullptr; // x, y, z = sb_vertex.uv of corner 0, 1, 2
    const struct TexHost* texHost = nullptr; // host emulation of the same lookups (tests/emul only)
};

// SAH constants of the BVH2 -> BVH8 cut (collapse_dp_node).  Ylitie 2017 uses node : triangle = 1 : 0.3; measured
// here on the 2 M- and 10 M-triangle scenes 0.8-1.2 is best (a triangle test is fewer instructions than a node
// test, but it runs with few lanes of the warp active).  With one segment per leaf slot the curve constant is moot.
#ifndef SB_COST_TRI
#define SB_COST_TRI 1.0f
#endif
constexpr float kCostNode = 1.0f, kCostTri = SB_COST_TRI, kCostSeg = 1.0f;
#ifndef SB_MAX_LEAF
#define SB_MAX_LEAF 3
#endif
constexpr uint32_t kMaxLeafTris = SB_MAX_LEAF; // triangles per leaf slot (the 24-bit leaf field of a node holds 8 x 3)

struct WideBvh
{
    WideNode* nodes = nullptr;
    uint32_t numNodes = 0;
    uint32_t* primOrder = nullptr; // leaf-order position -> index into the unsorted primitive list
    uint32_t numPrims = 0;
    uint32_t depth = 0; // levels of wide nodes (root = 1)
};

// double-precision inverse of an affine 3x4, rounded to float once (same formula as the oracle, so
// both sides hold bit-identical world_to_object matrices)
inline Affine invert_affine(const Affine& a)
{
    const double m00 = a.m[0], m01 = a.m[1], m02 = a.m[2], tx = a.m[3];
    const double m10 = a.m[4], m11 = a.m[5], m12 = a.m[6], ty = a.m[7];
    const double m20 = a.m[8], m21 = a.m[9], m22 = a.m[10], tz = a.m[11];
    const double c00 = m11 * m22 - m12 * m21, c01 = m12 * m20 - m10 * m22, c02 = m10 * m21 - m11 * m20;
    const double det = m00 * c00 + m01 * c01 + m02 * c02;
    const double id = 1.0 / det;
    const double r0 = c00 * id, r1 = (m02 * m21 - m01 * m22) * id, r2 = (m01 * m12 - m02 * m11) * id;
    const double r3 = c01 * id, r4 = (m00 * m22 - m02 * m20) * id, r5 = (m02 * m10 - m00 * m12) * id;
    const double r6 = c02 * id, r7 = (m01 * m20 - m00 * m21) * id, r8 = (m00 * m11 - m01 * m10) * id;
    Affine o;
    o.m[0] = float(r0);
    o.m[1] = float(r1);
    o.m[2] = float(r2);
    o.m[3] = float(-(r0 * tx + r1 * ty + r2 * tz));
    o.m[4] = float(r3);
    o.m[5] = float(r4);
    o.m[6] = float(r5);
    o.m[7] = float(-(r3 * tx + r4 * ty + r5 * tz));
    o.m[8] = float(r6);
    o.m[9] = float(r7);
    o.m[10] = float(r8);
    o.m[11] = float(-(r6 * tx + r7 * ty + r8 * tz));
    return o;
}
inline float affine_uniform_scale(const Affine& a)
{
    const double m00 = a.m[0], m01 = a.m[1], m02 = a.m[2];
    const double m10 = a.m[4], m11 = a.m[5], m12 = a.m[6];
    const double m20 = a.m[8], m21 = a.m[9], m22 = a.m[10];
    const double det = m00 * (m11 * m22 - m12 * m21) + m01 * (m12 * m20 - m10 * m22) + m02 * (m10 * m21 - m11 * m20);
    return float(std::cbrt(std::fabs(det)));
}

// index of the last element of the ascending array `a` that is <= v
SB_HD uint32_t upper_owner(const uint32_t* a, uint32_t n, uint32_t v)
{
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] <= v)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// Build a wide BVH over `n` boxes.  boxes are consumed (device memory, Aabb per primitive).
inline WideBvh build_wide_bvh(Exec& ex, const Aabb* boxes, uint32_t n, uint32_t maxLeaf, float cNode, float cPrim)
{
    WideBvh out;
    out.numPrims = n;
    if (n == 0)
        return out;

    // ---- scene bounds (of centres) via a tiny two-level reduction written as parallel_for ------
    const uint32_t kChunks = 1024;
    Aabb* partial = ex.alloc<Aabb>(kChunks);
    ex.pfor(kChunks, SB_LAMBDA(size_t c) {
        Aabb b = aabb_empty();
        for (size_t i = c; i < n; i += kChunks)
            aabb_grow(b, (boxes[i].lo + boxes[i].hi) * 0.5f);
        partial[c] = b;
    });
    Aabb* total = ex.alloc<Aabb>(1);
    ex.pfor(1, SB_LAMBDA(size_t) {
        Aabb b = aabb_empty();
        for (uint32_t c = 0; c < kChunks; ++c)
            b = aabb_union(b, partial[c]);
        total[0] = b;
    });

    ex.mark("  bounds");
    // ---- Morton codes + sort ---------------------------------------------------------------------
    uint64_t* keys = ex.alloc<uint64_t>(n);
    uint64_t* keysSorted = ex.alloc<uint64_t>(n);
    uint32_t* vals = ex.alloc<uint32_t>(n);
    uint32_t* sorted = ex.alloc<uint32_t>(n); // Morton position -> primitive index
    ex.pfor(n, SB_LAMBDA(size_t i) {
        const Aabb sb_ = total[0];
        const float3 ext = sb_.hi - sb_.lo;
        const float3 inv = mk3(ext.x > 0.0f ? 1.0f / ext.x : 0.0f, ext.y > 0.0f ? 1.0f / ext.y : 0.0f, ext.z > 0.0f ? 1.0f / ext.z : 0.0f);
        keys[i] = morton63((boxes[i].lo + boxes[i].hi) * 0.5f, sb_.lo, inv);
        vals[i] = uint32_t(i);
    });
    ex.sort_pairs_u64_u32(keys, keysSorted, vals, sorted, n);
    ex.free(keys);
    ex.free(keysSorted);
    ex.free(vals);
    ex.free(partial);
    ex.free(total);

    ex.mark("  morton + sort");
    // ---- PLOC ----------------------------------------------------------------------------------
    const uint32_t numNodes2 = 2 * n - 1;
    Bvh2Node* nodes = ex.alloc<Bvh2Node>(numNodes2);
    uint32_t* count = ex.alloc<uint32_t>(numNodes2);
    uint32_t* clusterA = ex.alloc<uint32_t>(n);
    uint32_t* clusterB = ex.alloc<uint32_t>(n);
    uint32_t* nearest = ex.alloc<uint32_t>(n);
    uint64_t* flags = ex.alloc<uint64_t>(n + 1);
    uint64_t* scan = ex.alloc<uint64_t>(n + 1);
    ex.pfor(n, SB_LAMBDA(size_t i) {
        Bvh2Node nd;
        node_set_box(nd, boxes[sorted[i]]);
        nd.left = kInvalid;
        nd.right = kInvalid;
        nodes[i] = nd;
        count[i] = 1u;
        clusterA[i] = uint32_t(i);
    });
    uint32_t active = n, nextNode = n;
    std::vector<std::pair<uint32_t, uint32_t>> rounds; // (first node, nodes) created by each PLOC round
    uint32_t* cur = clusterA;
    uint32_t* nxt = clusterB;
    while (active > 1)
    {
        const uint32_t na = active;
        const uint32_t* curC = cur;
        ex.pfor(na, SB_LAMBDA(size_t i) { nearest[i] = ploc_nearest(nodes, curC, na, uint32_t(i)); });
        // flags: high 32 bits = this cluster starts a merge (creates a node), low 32 bits = survives
        ex.pfor(na + 1, SB_LAMBDA(size_t i) {
            if (i == na)
            {
                flags[i] = 0;
                return;
            }
            const uint32_t j = nearest[i];
            const bool mutual = (j != kInvalid) && nearest[j] == uint32_t(i);
            const bool merges = mutual && uint32_t(i) < j;
            const bool dies = mutual && uint32_t(i) > j;
            flags[i] = (uint64_t(merges ? 1u : 0u) << 32) | uint64_t(dies ? 0u : 1u);
        });
        ex.exclusive_scan_u64(flags, scan, na + 1);
        uint32_t* nxtC = nxt;
        const uint32_t base = nextNode;
        ex.pfor(na, SB_LAMBDA(size_t i) {
            const uint64_t f = flags[i], s = scan[i];
            if (!(f & 0xffffffffull))
                return; // absorbed by its partner
            const uint32_t pos = uint32_t(s & 0xffffffffull);
            if (f >> 32)
            {
                const uint32_t id = base + uint32_t(s >> 32);
                const uint32_t l = curC[i], r = curC[nearest[i]];
                Bvh2Node nd;
                node_set_box(nd, aabb_union(node_box(nodes[l]), node_box(nodes[r])));
                nd.left = l;
                nd.right = r;
                nodes[id] = nd;
                count[id] = count[l] + count[r];
                nxtC[pos] = id;
            }
            else
            {
                nxtC[pos] = curC[i];
            }
        });
        const uint64_t totals = ex.read(scan + na);
        const uint32_t merges = uint32_t(totals >> 32);
        active = uint32_t(totals & 0xffffffffull);
        rounds.emplace_back(nextNode, merges);
        nextNode += merges;
        std::swap(cur, nxt);
        if (merges == 0)
            throw std::runtime_error("PLOC made no progress");
    }
    ex.mark("  PLOC rounds", (long long)rounds.size());
    const uint32_t root = ex.read(cur); // the last surviving cluster
    ex.free(clusterA);
    ex.free(clusterB);
    ex.free(nearest);

    // ---- optimal cut table, bottom-up along the merge rounds ------------------------------------------
    CollapseDp* dp = ex.alloc<CollapseDp>(n > 1 ? n - 1 : 1);
    for (const auto& rd : rounds)
    {
        const uint32_t first = rd.first;
        ex.pfor(rd.second, SB_LAMBDA(size_t i) { collapse_dp_node(nodes, count, dp, n, maxLeaf, cNode, cPrim, first + uint32_t(i)); });
    }

    ex.mark("  BVH8 cut DP");
    // ---- collapse to 8-wide, level by level --------------------------------------------------------
    // upper bound on wide nodes: every wide node except a degenerate root has >= 2 children and every
    // inner child holds > maxLeaf primitives -> fewer than n nodes; keep it simple and safe.
    const uint32_t maxWide = std::max<uint32_t>(1u, n);
    WideNode* wide = ex.alloc<WideNode>(maxWide);
    uint32_t* primOrder2 = ex.alloc<uint32_t>(n); // leaf-order position -> Morton position
    CollapseItem* itemsA = ex.alloc<CollapseItem>(maxWide);
    CollapseItem* itemsB = ex.alloc<CollapseItem>(maxWide);
    uint32_t* slotBuf = ex.alloc<uint32_t>(size_t(maxWide) * 8);
    CollapseItem rootItem;
    rootItem.bvh2Node = root;
    rootItem.wideIndex = 0;
    ex.write(itemsA, rootItem);
    uint32_t levelCount = 1, nodesUsed = 1, primsUsed = 0, levels = 0;
    CollapseItem* curI = itemsA;
    CollapseItem* nxtI = itemsB;
    while (levelCount > 0)
    {
        const uint32_t lc = levelCount;
        ++levels;
        const CollapseItem* items = curI;
        ex.pfor(lc + 1, SB_LAMBDA(size_t i) {
            if (i == lc)
            {
                flags[i] = 0;
                return;
            }
            uint32_t slots[8];
            uint32_t nPrims = 0;
            const uint32_t nInner = collapse_select(nodes, count, dp, n, items[i].bvh2Node, maxLeaf, slots, nPrims);
            for (int s = 0; s < 8; ++s)
                slotBuf[i * 8 + s] = slots[s];
            flags[i] = (uint64_t(nInner) << 32) | uint64_t(nPrims);
        });
        ex.exclusive_scan_u64(flags, scan, lc + 1);
        CollapseItem* nextItems = nxtI;
        const uint32_t nodeBase = nodesUsed, primBase = primsUsed;
        ex.pfor(lc, SB_LAMBDA(size_t i) {
            uint32_t slots[8];
            for (int s = 0; s < 8; ++s)
                slots[s] = slotBuf[i * 8 + s];
            const uint64_t s = scan[i];
            const uint32_t innerOff = uint32_t(s >> 32), primOff = uint32_t(s & 0xffffffffull);
            collapse_emit(nodes, count, n, items[i].bvh2Node, maxLeaf, slots, nodeBase + innerOff, primBase + primOff, wide[items[i].wideIndex],
                          nextItems, innerOff, primOrder2);
        });
        const uint64_t totals = ex.read(scan + lc);
        levelCount = uint32_t(totals >> 32);
        nodesUsed += levelCount;
        primsUsed += uint32_t(totals & 0xffffffffull);
        std::swap(curI, nxtI);
        if (nodesUsed > maxWide)
            throw std::runtime_error("wide BVH node estimate exceeded");
    }
    if (primsUsed != n)
        throw std::runtime_error("wide BVH lost primitives during collapse");

    ex.mark("  collapse levels", (long long)levels);
    // leaf order -> original primitive index
    uint32_t* primOrder = ex.alloc<uint32_t>(n);
    ex.pfor(n, SB_LAMBDA(size_t i) { primOrder[i] = sorted[primOrder2[i]]; });

    ex.free(primOrder2);
    ex.free(itemsA);
    ex.free(itemsB);
    ex.free(slotBuf);
    ex.free(flags);
    ex.free(scan);
    ex.free(dp);
    ex.free(nodes);
    ex.free(count);
    ex.free(sorted);

    ex.mark("  frees");
    out.nodes = wide;
    out.numNodes = nodesUsed;
    out.primOrder = primOrder;
    out.depth = levels;
    // The traversal stack holds at most one postponed node group per level above the current node, so a tree of
    // `levels` levels needs levels - 1 entries: checked HERE, once per build, instead of trusting a counter at run
    // time (an overflowing stack would silently drop geometry).
    if (levels > uint32_t(kStackSize))
        throw std::runtime_error("wide BVH is " + std::to_string(levels) + " levels deep; the traversal stack holds " +
                                 std::to_string(kStackSize) + " (degenerate geometry? raise kStackSize)");
    return out;
}

// Flatten instances to world space and build both BVHs.  `S` must already hold the uploaded scene
// arrays and the instance table (with firstPrim offsets per kind).  instTriFirst / instSegFirst are
// device arrays of length numInstances+1 with the exclusive prefix sums of triangles / segments per
// instance; segFirstPoint (per world segment) is precomputed on the host from the curve tables.
inline void build_scene_bvhs(Exec& ex, SceneDev& S, const uint32_t* instTriFirst, uint32_t numTris, const uint32_t* instSegFirst,
                             uint32_t numSegments, const SegInfo* segInfoUnsorted, uint32_t curveSplit)
{
    const uint32_t K = curveSplit ? curveSplit : 1u;
    const uint32_t numSegs = numSegments * K; // one traversal record per span
    S.numTris = numTris;
    S.numSegs = numSegs;
    const SceneDev Sv = S; // by-value copy for the lambdas
    const uint32_t numInst = S.numInstances;
    if (numTris)
    {
        TriRec* unsorted = ex.alloc<TriRec>(numTris);
        Aabb* boxes = ex.alloc<Aabb>(numTris);
        // shading-side record: ONE 64-byte fetch replaces the mesh -> index buffer -> 3 x vertex chain of the
        // reference's fillTriangleGeomData (closest_hit.cu:365-421): object-space corner positions, packed
        // normals and tangents, bit-for-bit the values the vertex buffer holds
        TriShade* shade = ex.alloc<TriShade>(numTris);
        ex.pfor(numTris, SB_LAMBDA(size_t g) {
            const uint32_t inst = upper_owner(instTriFirst, numInst + 1, uint32_t(g));
            const InstDev& I = Sv.instances[inst];
            const uint32_t t = uint32_t(g) - instTriFirst[inst];
            const sb_mesh m = Sv.meshes[I.geom];
            float3 p[3];
            uint32_t vi[3];
            for (int k = 0; k < 3; ++k)
            {
                vi[k] = m.vb_offset + Sv.indices[m.index + 3 * t + k];
                const sb_vertex& vx = Sv.vertices[vi[k]];
                p[k] = xform_point(I.o2w, mk3(vx.pos[0], vx.pos[1], vx.pos[2]));
            }
            TriShade sh;
            for (int k = 0; k < 3; ++k)
            {
                const sb_vertex& vx = Sv.vertices[vi[k]];
                sh.pos[3 * k] = vx.pos[0];
                sh.pos[3 * k + 1] = vx.pos[1];
                sh.pos[3 * k + 2] = vx.pos[2];
                sh.normal[k] = vx.normal;
                sh.tangent[k] = vx.tangent;
            }
            sh.material = I.material;
            shade[g] = sh;
            if (Sv.triUv)
            {
                uint4 puv;
                puv.x = Sv.vertices[vi[0]].uv;
                puv.y = Sv.vertices[vi[1]].uv;
                puv.z = Sv.vertices[vi[2]].uv;
                puv.w = 0u;
                Sv.triUv[g] = puv;
            }
            TriRec r;
            r.v0 = mk4(p[0], u2f(t));
            r.e1 = mk4(p[1] - p[0], u2f(inst | (I.mask << 28)));
            r.e2 = mk4(p[2] - p[0], u2f(uint32_t(g)));
            unsorted[g] = r;
            Aabb b = aabb_empty();
            aabb_grow(b, p[0]);
            aabb_grow(b, p[1]);
            aabb_grow(b, p[2]);
            boxes[g] = b;
        });
        ex.mark("flatten triangles", (long long)numTris);
        WideBvh bvh = build_wide_bvh(ex, boxes, numTris, kMaxLeafTris, kCostNode, kCostTri);
        TriRec* ordered = ex.alloc<TriRec>(numTris);
        const uint32_t* order = bvh.primOrder;
        ex.pfor(numTris, SB_LAMBDA(size_t i) { ordered[i] = unsorted[order[i]]; });
        ex.free(unsorted);
        ex.free(boxes);
        ex.free(bvh.primOrder);
        S.tris = ordered;
        S.triShade = shade;
        S.triNodes = bvh.nodes;
        S.numTriNodes = bvh.numNodes;
        S.triDepth = bvh.depth;
        ex.mark("reorder triangles");
    }
    if (numSegs)
    {
        SegRec* unsorted = ex.alloc<SegRec>(numSegs);
        Aabb* boxes = ex.alloc<Aabb>(numSegs);
        ex.pfor(numSegs, SB_LAMBDA(size_t g) {
            const SegInfo si = segInfoUnsorted[g / K];
            const uint32_t k = uint32_t(g % K);
            const InstDev& I = Sv.instances[si.inst];
            float4 q[4];
            for (int j = 0; j < 4; ++j)
            {
                const uint32_t pi = si.firstPoint + j;
                const float3 pw = xform_point(I.o2w, mk3(Sv.curvePoints[3 * pi], Sv.curvePoints[3 * pi + 1], Sv.curvePoints[3 * pi + 2]));
                const float rad = (pi < Sv.numCurveRadii ? Sv.curveRadii[pi] : 0.0f) * I.scale;
                q[j] = mk4(pw, rad);
            }
            const CurveSpan sp = curve_span(q, k, K);
            SegRec r;
            for (int j = 0; j < 4; ++j)
                r.q[j] = sp.c[j];
            Aabb b;
            curve_span_bounds(sp, b.lo, b.hi);
            unsorted[g] = r;
            boxes[g] = b;
        });
        ex.mark("flatten curve spans", (long long)numSegs);
        WideBvh bvh = build_wide_bvh(ex, boxes, numSegs, 1u, kCostNode, kCostSeg);
        SegRec* ordered = ex.alloc<SegRec>(numSegs);
        SegInfo* info = ex.alloc<SegInfo>(numSegs);
        const uint32_t* order = bvh.primOrder;
        ex.pfor(numSegs, SB_LAMBDA(size_t i) {
            const uint32_t g = order[i];
            ordered[i] = unsorted[g];
            SegInfo si = segInfoUnsorted[g / K];
            si.span = (g % K) | (K << 16);
            info[i] = si;
        });
        ex.free(unsorted);
        ex.free(boxes);
        ex.free(bvh.primOrder);
        S.segs = ordered;
        S.segInfo = info;
        S.segNodes = bvh.nodes;
        S.numSegNodes = bvh.numNodes;
        S.segDepth = bvh.depth;
        ex.mark("reorder curve spans");
    }
}

} // namespace sb
