// Compressed 8-wide BVH (CWBVH, after Ylitie, Karras, Laine 2017) -- data layout, the per-element
// steps of the on-device builder, and the traversal core.
//
// This replaces what the reference delegates to NVIDIA's closed OptiX driver: optixAccelBuild of one
// GAS per mesh/curve + one IAS (OptixRender.cpp:218-316,318-386,388-496) and optixTrace
// (OptixRender.cu:120-129, closest_hit.cu:185-197).  Design choice for 180 GB of HBM3e: instances are
// flattened to world space once (the reference's Hydra path already duplicates geometry per instance,
// RenderPass.cpp:252-257) and ONE wide BVH per primitive kind is built over all of it, so traversal
// never transforms rays and never chases an instance indirection.
//
// Node (80 B of content = 5 x 16 B; stored at a 96-byte stride, 32-byte aligned, and fetched with three 256-bit loads
// when SB_NODE96, else packed and fetched with five 128-bit loads):
//   n0 = { p.x, p.y, p.z, ex | ey<<8 | ez<<16 | imask<<24 }      quantisation frame + inner-node mask
//   n1 = { childBase, primBase, valid, 0 }      (SB_FIXED_BITS = 0: { childBase, primBase, meta[0..3], meta[4..7] })
//   n2 = { qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7] }
//   n3 = { qlo.z[0..3], qlo.z[4..7], qhi.x[0..3], qhi.x[4..7] }
//   n4 = { qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7] }
// child box = p + q * 2^(e-127).  valid = the bits of a 32-bit hit mask the node's slots own: slot s owns bits
// 3s..3s+2 if it is a leaf (one bit per primitive, unary; its first primitive is primBase + popc(valid below bit 3s))
// and bit 24+s if it is an inner node (its node is childBase + popc(imask below s)); an empty slot owns nothing.
// (meta, the encoding of Ylitie et al.: 0 empty | inner: 0x20 | (24 + slot) | leaf: unary(count)<<5 | first primitive
// offset relative to primBase, < 24.)  Slot s prefers children lying towards
// ((s&1?+:-), (s&2?+:-), (s&4?+:-)) of the node centre, which makes (slot ^ octant) a front-to-back order.
#pragma once
#include "hd.cuh"

#ifndef SB_FIXED_BITS
#define SB_FIXED_BITS 1 // hit-mask assembly of a node visit: see traverse.cuh
#endif

namespace sb
{

struct Aabb
{
    float3 lo, hi;
};
SB_HD Aabb aabb_empty()
{
    Aabb b;
    b.lo = mk3(3.0e38f);
    b.hi = mk3(-3.0e38f);
    return b;
}
SB_HD Aabb aabb_union(const Aabb& a, const Aabb& b)
{
    Aabb r;
    r.lo = mk3(fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z));
    r.hi = mk3(fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z));
    return r;
}
SB_HD void aabb_grow(Aabb& a, const float3& p)
{
    a.lo = mk3(fminf(a.lo.x, p.x), fminf(a.lo.y, p.y), fminf(a.lo.z, p.z));
    a.hi = mk3(fmaxf(a.hi.x, p.x), fmaxf(a.hi.y, p.y), fmaxf(a.hi.z, p.z));
}
SB_HD float aabb_half_area(const Aabb& a)
{
    const float dx = a.hi.x - a.lo.x, dy = a.hi.y - a.lo.y, dz = a.hi.z - a.lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// Binary BVH node used during construction.  Nodes [0, N) are the leaves (leaf i = i-th primitive in
// Morton order), nodes [N, 2N-1) are internal.
struct Bvh2Node
{
    float lo[3];
    uint32_t left;
    float hi[3];
    uint32_t right;
};
SB_HD Aabb node_box(const Bvh2Node& n)
{
    Aabb b;
    b.lo = mk3(n.lo[0], n.lo[1], n.lo[2]);
    b.hi = mk3(n.hi[0], n.hi[1], n.hi[2]);
    return b;
}
SB_HD void node_set_box(Bvh2Node& n, const Aabb& b)
{
    n.lo[0] = b.lo.x;
    n.lo[1] = b.lo.y;
    n.lo[2] = b.lo.z;
    n.hi[0] = b.hi.x;
    n.hi[1] = b.hi.y;
    n.hi[2] = b.hi.z;
}

// SB_NODE96: the 80-byte node padded to 96 bytes and 32-byte aligned, so that a visit fetches it with three 256-bit
// loads (LDG.E.ENL2.256, sm_100) instead of five 128-bit ones: with 32 divergent lanes every load instruction is 32 L1 tag
// look-ups, and the L1 data path is the busiest unit of the traversal kernels.
#ifndef SB_NODE96
#define SB_NODE96 1
#endif
#if SB_NODE96
struct alignas(32) WideNode
{
    uint4 n0, n1, n2, n3, n4, pad;
};
static_assert(sizeof(WideNode) == 96, "padded CWBVH node must be 96 bytes");
#else
struct WideNode
{
    uint4 n0, n1, n2, n3, n4;
};
static_assert(sizeof(WideNode) == 80, "CWBVH node must be 80 bytes");
#endif

// ---- 63-bit Morton code of a point in the unit cube ------------------------------------------------
SB_HD uint64_t spread21(uint64_t x)
{
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
SB_HD uint64_t morton63(const float3& p, const float3& lo, const float3& invExtent)
{
    const float fx = fminf(fmaxf((p.x - lo.x) * invExtent.x, 0.0f), 1.0f);
    const float fy = fminf(fmaxf((p.y - lo.y) * invExtent.y, 0.0f), 1.0f);
    const float fz = fminf(fmaxf((p.z - lo.z) * invExtent.z, 0.0f), 1.0f);
    const uint64_t x = uint64_t(fx * 2097151.0f), y = uint64_t(fy * 2097151.0f), z = uint64_t(fz * 2097151.0f);
    return (spread21(x) << 2) | (spread21(y) << 1) | spread21(z);
}

// ---- PLOC (Meister & Bittner 2018): nearest neighbour inside a window of the Morton-ordered clusters
#ifndef SB_PLOC_RADIUS
#define SB_PLOC_RADIUS 16
#endif
constexpr int kPlocRadius = SB_PLOC_RADIUS;

SB_HD uint32_t ploc_nearest(const Bvh2Node* nodes, const uint32_t* cluster, uint32_t n, uint32_t i)
{
    const Aabb bi = node_box(nodes[cluster[i]]);
    const uint32_t j0 = i > uint32_t(kPlocRadius) ? i - kPlocRadius : 0u;
    const uint32_t j1 = (i + kPlocRadius + 1 < n) ? i + kPlocRadius + 1 : n;
    float best = 3.0e38f;
    uint32_t bestJ = 0xffffffffu;
    for (uint32_t j = j0; j < j1; ++j)
    {
        if (j == i)
            continue;
        const float a = aabb_half_area(aabb_union(bi, node_box(nodes[cluster[j]])));
        if (a < best) // ties keep the lowest j: deterministic
        {
            best = a;
            bestJ = j;
        }
    }
    return bestJ;
}

// ---- BVH2 -> BVH8 collapse -----------------------------------------------------------------------
constexpr uint32_t kInvalid = 0xffffffffu;

struct CollapseItem
{
    uint32_t bvh2Node; // subtree root to turn into one wide node
    uint32_t wideIndex; // where the wide node goes
};

// Optimal BVH2 -> BVH8 cut by dynamic programming (Ylitie, Karras, Laine 2017, section 3.1).  For every BVH2
// node m and every i in 1..7, cost[i-1] is the SAH cost of the cheapest way to represent the subtree of m by a
// forest of at most i wide-node children (leaf slots or whole wide nodes):
//   C(m,1) = min( A(m) * P(m) * cPrim  if P(m) <= maxLeaf,      one leaf slot
//                 A(m) * cNode + min_k C(l,k) + C(r,8-k) )       a wide node of its own, split k : 8-k
//   C(m,i) = min( C(m,i-1), min_k C(l,k) + C(r,i-k) )           i >= 2
// A greedy "open the largest child" cut leaves the bottom of the tree badly filled: a balanced 32-triangle
// subtree becomes 8 wide nodes of 4 triangles each, where 7 leaf slots of 3 + one wide node of 11 do (the first
// build of the 2 M-triangle scene held 7 triangles per 80-byte node, and node tests are ~90 % of the traversal
// instructions).  The table is filled bottom-up along the PLOC merge rounds (children are always older nodes).
struct CollapseDp
{
    float cost[7];
    uint8_t split[8]; // [i-1], i = 2..7: k of the best split, 0 = "same as i-1"; [7]: k of the split as a wide node
};

SB_HD float collapse_cost(const Bvh2Node* nodes, const CollapseDp* dp, uint32_t numLeaves, float cPrim, uint32_t m, int i)
{
    if (m < numLeaves)
        return aabb_half_area(node_box(nodes[m])) * cPrim; // one primitive: one leaf slot, whatever the budget
    return dp[m - numLeaves].cost[i - 1];
}

// dp entry of inner BVH2 node m (both children already done)
SB_HD void collapse_dp_node(const Bvh2Node* nodes, const uint32_t* count, CollapseDp* dp, uint32_t numLeaves, uint32_t maxLeaf, float cNode,
                            float cPrim, uint32_t m)
{
    const uint32_t l = nodes[m].left, r = nodes[m].right;
    float cl[7], cr[7];
    for (int i = 1; i <= 7; ++i)
    {
        cl[i - 1] = collapse_cost(nodes, dp, numLeaves, cPrim, l, i);
        cr[i - 1] = collapse_cost(nodes, dp, numLeaves, cPrim, r, i);
    }
    const float area = aabb_half_area(node_box(nodes[m]));
    CollapseDp e;
    // as a wide node: 8 children to distribute
    float best = 3.0e38f;
    int bestK = 1;
    for (int k = 1; k <= 7; ++k)
    {
        const float c = cl[k - 1] + cr[8 - k - 1];
        if (c < best)
        {
            best = c;
            bestK = k;
        }
    }
    e.split[7] = uint8_t(bestK);
    const float cInternal = area * cNode + best;
    const float cLeaf = (count[m] <= maxLeaf) ? area * float(count[m]) * cPrim : 3.0e38f;
    e.cost[0] = fminf(cLeaf, cInternal);
    e.split[0] = 0;
    for (int i = 2; i <= 7; ++i)
    {
        float bd = 3.0e38f;
        int bk = 1;
        for (int k = 1; k < i; ++k)
        {
            const float c = cl[k - 1] + cr[i - k - 1];
            if (c < bd)
            {
                bd = c;
                bk = k;
            }
        }
        if (bd < e.cost[i - 2])
        {
            e.cost[i - 1] = bd;
            e.split[i - 1] = uint8_t(bk);
        }
        else
        {
            e.cost[i - 1] = e.cost[i - 2];
            e.split[i - 1] = 0;
        }
    }
    dp[m - numLeaves] = e;
}

// The (at most 8) children of the wide node rooted at BVH2 node `root`, following the dp decisions, then the
// octant-aware slot assignment.  `count[]` = primitives under each BVH2 node.  A child with more than maxLeaf
// primitives becomes a wide node of its own, any other a leaf slot.  Returns the number of inner children;
// nPrims = primitives referenced directly by this node's leaf slots.
SB_HD uint32_t collapse_select(const Bvh2Node* nodes, const uint32_t* count, const CollapseDp* dp, uint32_t numLeaves, uint32_t root,
                               uint32_t maxLeaf, uint32_t slots[8], uint32_t& nPrims)
{
    uint32_t cand[8];
    int n = 0;
    if (root < numLeaves)
    {
        cand[n++] = root; // a tree that is a single leaf
    }
    else
    {
        uint32_t stackNode[8];
        int stackBudget[8];
        int sp = 0;
        const int k8 = dp[root - numLeaves].split[7];
        stackNode[sp] = nodes[root].right;
        stackBudget[sp++] = 8 - k8;
        stackNode[sp] = nodes[root].left;
        stackBudget[sp++] = k8;
        while (sp > 0)
        {
            const uint32_t m = stackNode[--sp];
            int budget = stackBudget[sp];
            int k = 0;
            if (m >= numLeaves)
            {
                while (budget > 1 && (k = dp[m - numLeaves].split[budget - 1]) == 0)
                    --budget;
            }
            if (m < numLeaves || budget == 1)
            {
                cand[n++] = m; // at most 8 by construction of the budgets
                continue;
            }
            stackNode[sp] = nodes[m].right;
            stackBudget[sp++] = budget - k;
            stackNode[sp] = nodes[m].left;
            stackBudget[sp++] = k;
        }
    }
    // octant-aware assignment: greedily give the (child, slot) pair with the largest
    // dot(child centre - node centre, slot direction) until every child has a slot
    const Aabb pb = node_box(nodes[root]);
    const float3 pc = (pb.lo + pb.hi) * 0.5f;
    float3 off[8];
    for (int k = 0; k < n; ++k)
    {
        const Aabb cb = node_box(nodes[cand[k]]);
        off[k] = (cb.lo + cb.hi) * 0.5f - pc;
    }
    bool childDone[8] = { false, false, false, false, false, false, false, false };
    for (int s = 0; s < 8; ++s)
        slots[s] = kInvalid;
    for (int it = 0; it < n; ++it)
    {
        float bestCost = -3.0e38f;
        int bc = -1, bs = -1;
        for (int k = 0; k < n; ++k)
        {
            if (childDone[k])
                continue;
            for (int s = 0; s < 8; ++s)
            {
                if (slots[s] != kInvalid)
                    continue;
                const float cost = ((s & 1) ? off[k].x : -off[k].x) + ((s & 2) ? off[k].y : -off[k].y) + ((s & 4) ? off[k].z : -off[k].z);
                if (cost > bestCost)
                {
                    bestCost = cost;
                    bc = k;
                    bs = s;
                }
            }
        }
        childDone[bc] = true;
        slots[bs] = cand[bc];
    }
    uint32_t nInner = 0;
    nPrims = 0;
    for (int s = 0; s < 8; ++s)
    {
        const uint32_t c = slots[s];
        if (c == kInvalid)
            continue;
        if (c >= numLeaves && count[c] > maxLeaf)
            ++nInner;
        else
            nPrims += (c < numLeaves) ? 1u : count[c];
    }
    return nInner;
}

// leaves under a small subtree (<= 3), in left-to-right order
SB_HD uint32_t collect_leaves(const Bvh2Node* nodes, uint32_t numLeaves, uint32_t root, uint32_t out[4])
{
    uint32_t stack[8];
    int sp = 0;
    uint32_t n = 0;
    stack[sp++] = root;
    while (sp > 0)
    {
        const uint32_t c = stack[--sp];
        if (c < numLeaves)
        {
            if (n < 4)
                out[n++] = c;
        }
        else
        {
            if (sp + 2 <= 8)
            {
                stack[sp++] = nodes[c].right;
                stack[sp++] = nodes[c].left;
            }
        }
    }
    return n;
}

// biased exponent e (uint8) such that extent <= 255 * 2^(e-127)
SB_HD uint32_t quant_exponent(float extent)
{
    if (!(extent > 0.0f))
        return 1u; // degenerate axis: any tiny positive cell size works
    const float cell = extent / 255.0f;
    uint32_t e = (f2u(cell) + 0x7fffffu) >> 23; // round the magnitude up to a power of two
    if (e < 1u)
        e = 1u;
    if (e > 238u)
        e = 238u; // wide_node_hits scales the cell size by 2^15: stay clear of the float exponent range
    // keep one cell of head-room so that ceil() of the far face can never reach 256
    if (extent * (1.0f / u2f(e << 23)) > 254.0f)
        e += 1u;
    return e;
}

// Writes the compressed node for the slots chosen by collapse_select.  Emits the next level's work
// items (inner children, slot order) and this node's primitives (leaf slots, slot order).
SB_HD void collapse_emit(const Bvh2Node* nodes, const uint32_t* count, uint32_t numLeaves, uint32_t root, uint32_t maxLeaf,
                         const uint32_t slots[8], uint32_t childBase, uint32_t primBase, WideNode& out, CollapseItem* nextItems,
                         uint32_t nextOffset, uint32_t* primOrder)
{
    const Aabb pb = node_box(nodes[root]);
    const uint32_t ex = quant_exponent(pb.hi.x - pb.lo.x), ey = quant_exponent(pb.hi.y - pb.lo.y), ez = quant_exponent(pb.hi.z - pb.lo.z);
    const float sx = u2f(ex << 23), sy = u2f(ey << 23), sz = u2f(ez << 23); // cell sizes (powers of two)
    const float isx = 1.0f / sx, isy = 1.0f / sy, isz = 1.0f / sz;
    uint8_t meta[8], qlo[3][8], qhi[3][8];
    uint32_t imask = 0, innerRank = 0, primOff = 0, leafBits = 0;
    for (int s = 0; s < 8; ++s)
    {
        const uint32_t c = slots[s];
        if (c == kInvalid)
        {
            meta[s] = 0;
            for (int a = 0; a < 3; ++a)
            {
                qlo[a][s] = 255; // inverted box: can never be hit
                qhi[a][s] = 0;
            }
            continue;
        }
        const Aabb cb = node_box(nodes[c]);
        // conservative quantisation (power-of-two cells make the products exact)
        const float lo[3] = { floorf((cb.lo.x - pb.lo.x) * isx), floorf((cb.lo.y - pb.lo.y) * isy), floorf((cb.lo.z - pb.lo.z) * isz) };
        const float hi[3] = { ceilf((cb.hi.x - pb.lo.x) * isx), ceilf((cb.hi.y - pb.lo.y) * isy), ceilf((cb.hi.z - pb.lo.z) * isz) };
        const float plo[3] = { pb.lo.x, pb.lo.y, pb.lo.z };
        const float cs[3] = { sx, sy, sz };
        const float clo[3] = { cb.lo.x, cb.lo.y, cb.lo.z };
        const float chi[3] = { cb.hi.x, cb.hi.y, cb.hi.z };
        for (int a = 0; a < 3; ++a)
        {
            float l = fminf(fmaxf(lo[a], 0.0f), 255.0f), h = fminf(fmaxf(hi[a], 0.0f), 255.0f);
            // make sure the dequantised box really contains the child box (subtraction above rounds)
            while (l > 0.0f && plo[a] + l * cs[a] > clo[a])
                l -= 1.0f;
            while (h < 255.0f && plo[a] + h * cs[a] < chi[a])
                h += 1.0f;
            qlo[a][s] = uint8_t(l);
            qhi[a][s] = uint8_t(h);
        }
        if (c >= numLeaves && count[c] > maxLeaf)
        {
            meta[s] = uint8_t(0x20u | (24u + uint32_t(s)));
            imask |= 1u << s;
            CollapseItem it;
            it.bvh2Node = c;
            it.wideIndex = childBase + innerRank;
            nextItems[nextOffset + innerRank] = it;
            ++innerRank;
        }
        else
        {
            uint32_t leaves[4];
            const uint32_t nl = collect_leaves(nodes, numLeaves, c, leaves);
            for (uint32_t k = 0; k < nl; ++k)
                primOrder[primBase + primOff + k] = leaves[k];
            const uint32_t unary = (nl >= 3) ? 7u : ((nl == 2) ? 3u : 1u);
            meta[s] = uint8_t((unary << 5) | primOff);
            leafBits |= unary << (3 * s);
            primOff += nl;
        }
    }
    auto pack4 = [](const uint8_t* b) { return uint32_t(b[0]) | (uint32_t(b[1]) << 8) | (uint32_t(b[2]) << 16) | (uint32_t(b[3]) << 24); };
    out.n0.x = f2u(pb.lo.x);
    out.n0.y = f2u(pb.lo.y);
    out.n0.z = f2u(pb.lo.z);
    out.n0.w = ex | (ey << 8) | (ez << 16) | (imask << 24);
    out.n1.x = childBase;
    out.n1.y = primBase;
#if SB_FIXED_BITS
    out.n1.z = leafBits | (imask << 24); // the bits of a hit mask this node's slots own (traverse.cuh)
    out.n1.w = 0u;
    (void)meta;
#else
    out.n1.z = pack4(meta);
    out.n1.w = pack4(meta + 4);
#endif
    out.n2.x = pack4(qlo[0]);
    out.n2.y = pack4(qlo[0] + 4);
    out.n2.z = pack4(qlo[1]);
    out.n2.w = pack4(qlo[1] + 4);
    out.n3.x = pack4(qlo[2]);
    out.n3.y = pack4(qlo[2] + 4);
    out.n3.z = pack4(qhi[0]);
    out.n3.w = pack4(qhi[0] + 4);
    out.n4.x = pack4(qhi[1]);
    out.n4.y = pack4(qhi[1] + 4);
    out.n4.z = pack4(qhi[2]);
    out.n4.w = pack4(qhi[2] + 4);
}

} // namespace sb
