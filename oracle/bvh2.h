// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Plain binary BVH (binned SAH, top-down) over axis-aligned boxes, with a stack traversal.
// Stands in for optixAccelBuild / optixTrace (OptixRender.cpp:300-308,366-376,487-492;
// OptixRender.cu:120-129), which are closed NVIDIA driver code -> "parity unpinned" for the
// acceleration structure itself; hits are made structure-independent by an exact tie-break
// (equal t -> lower primitive id wins) that the CUDA traversal implements as well.
#pragma once
#include "vec.h"
#include <atomic>
#include <thread>
#include <vector>
#include <numeric>
#include <limits>

namespace orc
{

struct Aabb
{
    f3 lo, hi;
    void reset()
    {
        lo = mk3(std::numeric_limits<float>::infinity());
        hi = mk3(-std::numeric_limits<float>::infinity());
    }
    void grow(const f3& p)
    {
        lo = f3{ std::fmin(lo.x, p.x), std::fmin(lo.y, p.y), std::fmin(lo.z, p.z) };
        hi = f3{ std::fmax(hi.x, p.x), std::fmax(hi.y, p.y), std::fmax(hi.z, p.z) };
    }
    void grow(const Aabb& b)
    {
        grow(b.lo);
        grow(b.hi);
    }
    float half_area() const
    {
        const f3 d = hi - lo;
        return d.x * d.y + d.y * d.z + d.z * d.x;
    }
};

struct Bvh2Node
{
    Aabb box;
    uint32_t left; // internal: index of left child (right = left+1); leaf: first prim slot
    uint32_t count; // 0 = internal, else number of prims
};

struct Bvh2
{
    std::vector<Bvh2Node> nodes;
    std::vector<uint32_t> prims; // permutation of primitive ids

    struct Task
    {
        uint32_t node, first, count;
    };

    // Top-down binned-SAH construction of the subtrees in `stack` into `out` (children appended at out.size()).
    // Tasks of at most `deferBelow` primitives are not built but moved to `deferred` (parallel build, below).
    void build_tasks(std::vector<Bvh2Node>& out, std::vector<Task>& stack, const std::vector<Aabb>& boxes, const std::vector<f3>& cent,
                     uint32_t deferBelow, std::vector<Task>* deferred)
    {
        auto& nodes = out;
        constexpr int kBins = 16;
        while (!stack.empty())
        {
            const Task t = stack.back();
            stack.pop_back();
            if (deferred && t.count <= deferBelow)
            {
                deferred->push_back(t);
                continue;
            }
            Aabb box, cbox;
            box.reset();
            cbox.reset();
            for (uint32_t i = t.first; i < t.first + t.count; ++i)
            {
                box.grow(boxes[prims[i]]);
                cbox.grow(cent[prims[i]]);
            }
            nodes[t.node].box = box;
            if (t.count <= 2)
            {
                nodes[t.node].left = t.first;
                nodes[t.node].count = t.count;
                continue;
            }
            // pick best (axis, bin) split
            float bestCost = std::numeric_limits<float>::infinity();
            int bestAxis = -1, bestBin = -1;
            for (int axis = 0; axis < 3; ++axis)
            {
                const float lo = (&cbox.lo.x)[axis], hi = (&cbox.hi.x)[axis];
                if (!(hi > lo))
                    continue;
                const float scale = kBins / (hi - lo);
                Aabb bb[kBins];
                uint32_t bc[kBins] = {};
                for (auto& b : bb)
                    b.reset();
                for (uint32_t i = t.first; i < t.first + t.count; ++i)
                {
                    const uint32_t p = prims[i];
                    int b = int(((&cent[p].x)[axis] - lo) * scale);
                    b = std::min(std::max(b, 0), kBins - 1);
                    bb[b].grow(boxes[p]);
                    bc[b]++;
                }
                float rightArea[kBins];
                uint32_t rightCnt[kBins];
                Aabb acc;
                acc.reset();
                uint32_t cnt = 0;
                for (int b = kBins - 1; b > 0; --b)
                {
                    acc.grow(bb[b]);
                    cnt += bc[b];
                    rightArea[b] = cnt ? acc.half_area() : 0.0f;
                    rightCnt[b] = cnt;
                }
                acc.reset();
                cnt = 0;
                for (int b = 0; b < kBins - 1; ++b)
                {
                    acc.grow(bb[b]);
                    cnt += bc[b];
                    if (cnt == 0 || rightCnt[b + 1] == 0)
                        continue;
                    const float cost = acc.half_area() * cnt + rightArea[b + 1] * rightCnt[b + 1];
                    if (cost < bestCost)
                    {
                        bestCost = cost;
                        bestAxis = axis;
                        bestBin = b;
                    }
                }
            }
            const float leafCost = box.half_area() * t.count;
            const bool makeLeaf = (bestAxis < 0) ? (t.count <= 8) : (t.count <= 4 && bestCost >= leafCost);
            if (makeLeaf)
            {
                nodes[t.node].left = t.first;
                nodes[t.node].count = t.count;
                continue;
            }
            uint32_t mid;
            if (bestAxis < 0)
            {
                mid = t.first + t.count / 2; // all centroids coincide: median split
            }
            else
            {
                const float lo = (&cbox.lo.x)[bestAxis], hi = (&cbox.hi.x)[bestAxis];
                const float scale = kBins / (hi - lo);
                uint32_t i = t.first, j = t.first + t.count;
                while (i < j)
                {
                    int b = int(((&cent[prims[i]].x)[bestAxis] - lo) * scale);
                    b = std::min(std::max(b, 0), kBins - 1);
                    if (b <= bestBin)
                        ++i;
                    else
                        std::swap(prims[i], prims[--j]);
                }
                mid = i;
                if (mid == t.first || mid == t.first + t.count)
                    mid = t.first + t.count / 2;
            }
            const uint32_t l = uint32_t(nodes.size());
            nodes.push_back(Bvh2Node{});
            nodes.push_back(Bvh2Node{});
            nodes[t.node].left = l;
            nodes[t.node].count = 0;
            stack.push_back(Task{ l, t.first, mid - t.first });
            stack.push_back(Task{ l + 1, mid, t.first + t.count - mid });
        }
    }

    // The tree is built top-down; subtrees are independent, so the top of the tree is built serially until the open
    // subtrees are small, then those are built on all host threads into private node arrays and appended.  Hits do
    // not depend on the node order (exact tie-break on the primitive id), only the build time does.
    void build(const std::vector<Aabb>& boxes)
    {
        const uint32_t n = uint32_t(boxes.size());
        nodes.clear();
        prims.resize(n);
        std::iota(prims.begin(), prims.end(), 0u);
        if (n == 0)
            return;
        std::vector<f3> cent(n);
        for (uint32_t i = 0; i < n; ++i)
            cent[i] = (boxes[i].lo + boxes[i].hi) * 0.5f;
        nodes.reserve(2 * size_t(n));
        nodes.push_back(Bvh2Node{});
        std::vector<Task> stack, deferred;
        stack.push_back(Task{ 0, 0, n });
        unsigned nthreads = std::thread::hardware_concurrency();
        if (nthreads == 0)
            nthreads = 1;
        if (n < 50000u || nthreads == 1)
        {
            build_tasks(nodes, stack, boxes, cent, 0u, nullptr);
            return;
        }
        build_tasks(nodes, stack, boxes, cent, std::max<uint32_t>(n / (8u * nthreads), 1024u), &deferred);
        std::vector<std::vector<Bvh2Node>> sub(deferred.size());
        std::atomic<size_t> next{ 0 };
        auto work = [&]() {
            for (;;)
            {
                const size_t k = next.fetch_add(1);
                if (k >= deferred.size())
                    return;
                std::vector<Bvh2Node>& loc = sub[k];
                loc.reserve(2 * size_t(deferred[k].count));
                loc.push_back(Bvh2Node{});
                std::vector<Task> st;
                st.push_back(Task{ 0, deferred[k].first, deferred[k].count });
                build_tasks(loc, st, boxes, cent, 0u, nullptr);
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nthreads; ++t)
            pool.emplace_back(work);
        work();
        for (auto& th : pool)
            th.join();
        for (size_t k = 0; k < deferred.size(); ++k)
        {
            // sub[k][0] is the subtree root (already allocated as nodes[deferred[k].node]); sub[k][i >= 1] lands at base + i - 1
            const uint32_t base = uint32_t(nodes.size());
            auto fix = [&](Bvh2Node nd) {
                if (nd.count == 0)
                    nd.left += base - 1u;
                return nd;
            };
            nodes[deferred[k].node] = fix(sub[k][0]);
            for (size_t i = 1; i < sub[k].size(); ++i)
                nodes.push_back(fix(sub[k][i]));
            sub[k] = std::vector<Bvh2Node>();
        }
    }

    // Slab test with the conservative tmax scale of Ize 2013 (1 + 2*gamma(3)); returns the entry
    // distance or +inf on a miss.
    static float hit_box(const Aabb& b, const f3& o, const f3& inv, float tmin, float tmax)
    {
        float t0 = (b.lo.x - o.x) * inv.x, t1 = (b.hi.x - o.x) * inv.x;
        float tn = std::fmin(t0, t1), tf = std::fmax(t0, t1);
        t0 = (b.lo.y - o.y) * inv.y;
        t1 = (b.hi.y - o.y) * inv.y;
        tn = std::fmax(tn, std::fmin(t0, t1));
        tf = std::fmin(tf, std::fmax(t0, t1));
        t0 = (b.lo.z - o.z) * inv.z;
        t1 = (b.hi.z - o.z) * inv.z;
        tn = std::fmax(tn, std::fmin(t0, t1));
        tf = std::fmin(tf, std::fmax(t0, t1));
        tf *= 1.0000004f;
        tn = std::fmax(tn, tmin);
        return (tn <= std::fmin(tf, tmax)) ? tn : std::numeric_limits<float>::infinity();
    }

    // visit(primId, tmax&) -> true to stop traversal (any-hit); it may shrink tmax
    template <typename F>
    void traverse(const f3& o, const f3& d, float tmin, float tmax, F&& visit) const
    {
        if (nodes.empty())
            return;
        const f3 inv{ 1.0f / d.x, 1.0f / d.y, 1.0f / d.z };
        const float inf = std::numeric_limits<float>::infinity();
        if (hit_box(nodes[0].box, o, inv, tmin, tmax) == inf)
            return;
        struct Entry
        {
            uint32_t node;
            float tn;
        };
        Entry stack[128];
        int sp = 0;
        stack[sp++] = Entry{ 0, 0.0f };
        while (sp)
        {
            const Entry e = stack[--sp];
            if (e.tn > tmax)
                continue;
            const Bvh2Node& n = nodes[e.node];
            if (n.count)
            {
                for (uint32_t i = 0; i < n.count; ++i)
                {
                    if (visit(prims[n.left + i], tmax))
                        return;
                }
                continue;
            }
            const float ta = hit_box(nodes[n.left].box, o, inv, tmin, tmax);
            const float tb = hit_box(nodes[n.left + 1].box, o, inv, tmin, tmax);
            if (sp + 2 > 128)
                continue;
            if (ta <= tb)
            {
                if (tb != inf)
                    stack[sp++] = Entry{ n.left + 1, tb };
                if (ta != inf)
                    stack[sp++] = Entry{ n.left, ta };
            }
            else
            {
                if (ta != inf)
                    stack[sp++] = Entry{ n.left, ta };
                stack[sp++] = Entry{ n.left + 1, tb };
            }
        }
    }
};

} // namespace orc
