// Hand-written sm_100a kernels of the wavefront path tracer + their launchers.
//
// Launch geometry (B200: 148 SMs): every stage is a grid-stride kernel over a device-resident queue count,
// launched with a fixed grid of 148 * k blocks so that (a) no host round-trip is needed to size a launch, (b) the
// grid is always a whole number of SM-loads.  k = 8 (one resident wave at 64 registers) for the persistent
// traversal kernels, which balance themselves through their dynamic ray fetch; k = 128-256 for the statically
// partitioned kernels, where the block scheduler does the balancing (SB_*_GRID below).  Blocks are 128 threads:
// traversal and shading are bound by instruction issue and latency, so registers (occupancy) matter more than
// block size.
#include <cstdint>
namespace sb
{
__constant__ uint32_t c_sobol[5][32];
uint32_t h_sobol[5][32];
} // namespace sb

#include "wavefront.cuh"
#include "kernels.h"
#include <type_traits>

namespace sb
{

uint32_t* upload_sobol_table(cudaStream_t stream)
{
    sobol_generate(h_sobol);
    SB_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_sobol, h_sobol, sizeof(h_sobol), 0, cudaMemcpyHostToDevice, stream));
    static uint32_t tab[kSobolTabWords];
    sobol_build_tables(h_sobol, tab);
    uint32_t* d = nullptr;
    SB_CUDA_CHECK(cudaMalloc(&d, sizeof(tab)));
    SB_CUDA_CHECK(cudaMemcpyAsync(d, tab, sizeof(tab), cudaMemcpyHostToDevice, stream));
    SB_CUDA_CHECK(cudaStreamSynchronize(stream));
    return d;
}

constexpr int kBlock = 128;
constexpr int kBlockWarps = kBlock / 32;
#ifndef SB_TINY_NODES
#define SB_TINY_NODES 64
#endif
constexpr uint32_t kTinyBvhNodes = SB_TINY_NODES; // at or below this many wide nodes traversal is a few steps: no dynamic fetch

// Block-aggregated queue allocation.  The first ncu source view of k_shade had a third of its stall samples
// on the two warp-aggregated atomicAdds of the queue cursors: every warp of the chip hits the same two L2
// addresses once per iteration and same-address atomics serialise in the L2 slice.  Here all threads of a block
// reach the allocation point together (the callers loop block-uniformly), count with ballots, and ONE thread per
// queue issues the atomic for the whole block: 4x fewer atomics, and the survivors of a block land in
// consecutive slots.  `buf` alternates between consecutive calls so that one __syncthreads pair per call suffices.
struct BlockAlloc
{
    uint32_t warpCount[2][2][kBlockWarps]; // [buf][queue][warp]
    uint32_t base[2][2]; // [buf][queue]
};
// wantA / wantB: this thread needs a slot in queue A / B.  Returns the slots through slotA / slotB.
__device__ __forceinline__ void block_alloc2(BlockAlloc& sh, uint32_t buf, uint32_t* counterA, uint32_t* counterB, bool wantA, bool wantB,
                                             uint32_t& slotA, uint32_t& slotB)
{
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned maskA = __ballot_sync(0xffffffffu, wantA), maskB = __ballot_sync(0xffffffffu, wantB);
    if (lane == 0u)
    {
        sh.warpCount[buf][0][warp] = __popc(maskA);
        sh.warpCount[buf][1][warp] = __popc(maskB);
    }
    __syncthreads();
    if (threadIdx.x < 2u)
    {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < kBlockWarps; ++w)
            total += sh.warpCount[buf][threadIdx.x][w];
        uint32_t* counter = threadIdx.x == 0u ? counterA : counterB;
        sh.base[buf][threadIdx.x] = (total != 0u && counter != nullptr) ? atomicAdd(counter, total) : 0u;
    }
    __syncthreads();
    uint32_t offA = sh.base[buf][0], offB = sh.base[buf][1];
#pragma unroll
    for (int w = 0; w < kBlockWarps; ++w)
    {
        if (w < int(warp))
        {
            offA += sh.warpCount[buf][0][w];
            offB += sh.warpCount[buf][1][w];
        }
    }
    const unsigned lt = (1u << lane) - 1u;
    slotA = offA + __popc(maskA & lt);
    slotB = offB + __popc(maskB & lt);
}

__global__ void __launch_bounds__(kBlock) k_raygen(FrameParams P, Queues Q)
{
    __shared__ BlockAlloc s_alloc;
    const uint32_t n = P.nPixPadded * P.chunk;
    uint32_t buf = 0;
    for (uint32_t base = blockIdx.x * kBlock; base < n; base += gridDim.x * kBlock, buf ^= 1u)
    {
        const uint32_t i = base + threadIdx.x;
        PathState ps;
        bool valid = false;
        if (i < n)
        {
            Q.Lacc[i] = mk4(0.0f, 0.0f, 0.0f, 0.0f);
            valid = raygen_state(P, i, ps);
        }
        uint32_t slot, unused;
        block_alloc2(s_alloc, buf, &Q.counts[0], nullptr, valid, false, slot, unused);
        if (valid)
        {
            Q.rayO[0][slot] = mk4(ps.o, u2f(i));
            Q.rayD[0][slot] = mk4(ps.d, 0.0f);
            Q.thr[0][slot] = mk4(1.0f, 1.0f, 1.0f, u2f(0u));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&Q.stats->paths, (unsigned long long)(P.width) * P.height * P.chunk);
}

// flush per-thread traversal statistics with one atomic per warp
__device__ __forceinline__ void flush_stats(StatCounters* g, const TravStats& st, bool shadow)
{
    unsigned n = st.nodes, t = st.tris, s = st.segs, o = st.overflow;
    for (int off = 16; off > 0; off >>= 1)
    {
        n += __shfl_xor_sync(0xffffffffu, n, off);
        t += __shfl_xor_sync(0xffffffffu, t, off);
        s += __shfl_xor_sync(0xffffffffu, s, off);
        o += __shfl_xor_sync(0xffffffffu, o, off);
    }
    if ((threadIdx.x & 31) == 0)
    {
        if (n)
            atomicAdd(shadow ? &g->nodesSh : &g->nodes, (unsigned long long)n);
        if (t)
            atomicAdd(shadow ? &g->trisSh : &g->tris, (unsigned long long)t);
        if (s)
            atomicAdd(shadow ? &g->segsSh : &g->segs, (unsigned long long)s);
        if (o)
            atomicAdd(&g->overflow, (unsigned long long)o);
    }
}

// Camera rays generated and traced in one kernel (images whose size is a whole number of 8x4 tiles: no padding
// pixels, so path id == queue slot and nothing needs allocating).  Saves writing the primary rays only to read them
// back, and one launch per batch.
template <bool STATS, bool CURVES>
__global__ void __launch_bounds__(kBlock, 8) k_primary(FrameParams P, SceneDev S, Queues Q)
{
    const uint32_t n = P.nPixPadded * P.chunk;
    TravStats st = { 0, 0, 0, 0 };
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        PathState ps;
        raygen_state(P, i, ps);
        float4 ha;
        uint32_t hb;
        trace_closest<STATS, CURVES>(S, ps.o, ps.d, P.materialTmin, ha, hb, &st);
        Q.Lacc[i] = mk4(0.0f, 0.0f, 0.0f, 0.0f);
        Q.rayO[0][i] = mk4(ps.o, u2f(i));
        Q.rayD[0][i] = mk4(ps.d, 0.0f);
        Q.thr[0][i] = mk4(1.0f, 1.0f, 1.0f, u2f(0u));
        Q.hitA[i] = ha;
        Q.hitB[i] = hb;
    }
    if (STATS)
        flush_stats(Q.stats, st, false);
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        Q.counts[count_path(0)] = n;
        atomicAdd(&Q.stats->paths, (unsigned long long)n);
        atomicAdd(&Q.stats->radianceRays, (unsigned long long)n);
    }
}

// ---- persistent traversal kernels with dynamic ray fetch ----------------------------------------------
// Incoherent secondary rays finish after very different numbers of steps; with one ray per thread the first
// ncu capture on the 2 M-triangle scene showed 7 of 32 lanes active on average.  Here every warp keeps its
// lanes busy: a lane whose ray is done pulls the next ray from the queue (one warp-aggregated atomic), and
// refills happen whenever fewer than kRefill lanes still hold a ray (Aila & Laine 2009 style, using
// trav_step so that node visits and primitive tests of different lanes interleave).
// Measured on the 2 M-triangle scene (profiles/r01_b_*): (1) batching primitive tests across the warp
// ("primitive postponing") did not pay off: the extra convergence points stop independent thread scheduling
// from interleaving the divergent node/primitive groups while they wait on loads; (2) the loop is extremely
// sensitive to its control-flow shape -- one extra early-return inside trav_node cost 50 % -- so keep it flat:
// one trav_step per lane, ONE warp-wide ballot per iteration.
#ifndef SB_REFILL
#define SB_REFILL 28
#endif
constexpr int kRefill = SB_REFILL;
#ifndef SB_PREFETCH_TRI_MB
#define SB_PREFETCH_TRI_MB 120 // B200: 126 MB of L2
#endif
constexpr size_t kPrefetchTriBytes = size_t(SB_PREFETCH_TRI_MB) << 20;

// Queue slots are handed to the warps in chunks (one L2 atomic per kFetchChunk rays instead of one per refill); the warp
// that takes a chunk prefetches its ray records, which the shade kernel streamed out to HBM, so that the refills that
// follow find them on chip instead of stalling the whole warp on a DRAM round trip (ncu: 5 % of all stall samples of the
// closest-hit kernel sat on the first use of a freshly fetched ray).
#ifndef SB_FETCH_CHUNK
#define SB_FETCH_CHUNK 32
#endif
constexpr uint32_t kFetchChunk = SB_FETCH_CHUNK;
struct WarpFetch
{
    uint32_t next = 0, end = 0; // warp-uniform: the slots this warp still owns
    bool exhausted = false; // warp-uniform: the queue has been handed out completely
};
__device__ __forceinline__ void prefetch_l2(const void* p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// hands out queue slots to the lanes with need == true; returns the slot or 0xffffffff.
// recA / recB (/ recC): the float4 record arrays of the queue, prefetched per chunk.
__device__ __forceinline__ uint32_t fetch_slots(WarpFetch& wf, uint32_t* head, uint32_t n, bool need, const float4* recA, const float4* recB,
                                                const float4* recC)
{
    const unsigned lane = threadIdx.x & 31u;
    unsigned mask = __ballot_sync(0xffffffffu, need);
    uint32_t slot = 0xffffffffu;
    while (mask != 0u)
    {
        if (wf.next == wf.end)
        {
            if (wf.exhausted)
                break;
            // chunk size: large for long queues, down to one warp-load for short ones (a short queue cut into 128-ray
            // chunks leaves some warps with two chunks and others with one: measured +7 % on the hair scene's shadow rays)
            const uint32_t warpsInGrid = gridDim.x * (blockDim.x >> 5);
            const uint32_t chunk = min(kFetchChunk, max(32u, (n / (warpsInGrid * 8u)) & ~31u));
            uint32_t base = 0;
            if (lane == 0u)
                base = atomicAdd(head, chunk);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= n)
            {
                wf.exhausted = true;
                break;
            }
            wf.next = base;
            wf.end = min(base + chunk, n);
            if (wf.end == n)
                wf.exhausted = true; // nothing beyond this chunk
#ifndef SB_FETCH_PREFETCH
#define SB_FETCH_PREFETCH 1
#endif
#if SB_FETCH_PREFETCH
            // 16-byte records: 8 per 128-byte line; every lane touches one line of each array
            for (uint32_t i = wf.next + lane * 8u; i < wf.end; i += 256u)
            {
                prefetch_l2(recA + i);
                prefetch_l2(recB + i);
                if (recC)
                    prefetch_l2(recC + i);
            }
#endif
        }
        const uint32_t avail = wf.end - wf.next;
        const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
        if (need && slot == 0xffffffffu && rank < avail)
            slot = wf.next + rank;
        wf.next += min(avail, uint32_t(__popc(mask)));
        mask = __ballot_sync(0xffffffffu, need && slot == 0xffffffffu);
    }
    return slot;
}

// ---- deferred primitive tests ---------------------------------------------------------------------------------
// In the plain loop a lane tests a primitive as soon as a node visit queues one, so primitive tests run with whatever
// few lanes happen to hold one (ncu: 4.7 of 32 lanes in the curve test of the hair scene, ~8 in the triangle test) while
// the rest of the warp waits.  Here a lane may park up to D leaf groups in shared memory and keep visiting nodes;
// the warp runs ONE primitive test per lane with pending work when at least THRESH lanes have some, or when a
// lane cannot go on otherwise (its queue is full, or it has no nodes left).  Results do not depend on the order of the
// tests (closest hit: ties break on the primitive id; any hit: a boolean).
// Depth of the parking queue and vote threshold per kernel family, measured (profiles/r02_ab_experiments.txt, r2y/r2z):
// curve scenes gain 5 % (closest hit) and 15 % (any hit); triangle-only scenes 1-2 % for closest-hit rays and nothing
// for any-hit rays.  0 = the plain loop.  (Prefetching the first parked primitive at the node visit: no gain on the
// hair scene, +35 % closest-hit time on the 2 M-triangle scene.)
#ifndef SB_DEFER_CURVES
#define SB_DEFER_CURVES 3
#endif
#ifndef SB_DEFER_CURVES_THRESH
#define SB_DEFER_CURVES_THRESH 16
#endif
#ifndef SB_DEFER_TRIS
#define SB_DEFER_TRIS 2
#endif
#ifndef SB_DEFER_TRIS_SHADOW
#define SB_DEFER_TRIS_SHADOW 0
#endif
#ifndef SB_DEFER_TRIS_THRESH
#define SB_DEFER_TRIS_THRESH 12
#endif
// all 32 lanes call this (it votes); returns whether the lane's traversal of the current BVH goes on
template <bool ANY, bool STATS, bool CURVES, bool PF, int XU, int D, int THRESH>
__device__ __forceinline__ bool trav_iter_defer(bool active, int phase, const SceneDev& S, Traversal& T, TravStack& K, uint32_t* dq, int& dc,
                                                uint32_t rayMask, Ray& ray, const RayPrep& rp, HitRec& hit, bool& anyHit, TravStats* st)
{
    const bool curvePhase = CURVES && phase == 1;
    bool pend = false, stuck = false, haveNodes = false;
    if (active)
    {
        haveNodes = T.ngroup.y > 0x00ffffffu || T.sp > 0;
        if (haveNodes && (T.tgroup.y == 0u || dc < D))
        {
            if (T.tgroup.y != 0u)
            {
                uint32_t* e = dq + dc * 3 * kBlock;
                e[0] = T.tgroup.x;
                e[kBlock] = T.tgroup.y;
                e[2 * kBlock] = T.tvalid;
                ++dc;
                T.tgroup.y = 0u;
            }
            trav_node<STATS, (SB_SMEM_STACK > 0), XU>(T, K, curvePhase ? S.segNodes : S.triNodes, ray, rp, st);
            haveNodes = T.ngroup.y > 0x00ffffffu || T.sp > 0;
        }
        pend = T.tgroup.y != 0u || dc > 0;
        stuck = pend && (!haveNodes || (T.tgroup.y != 0u && dc == D));
    }
    const unsigned waiting = __ballot_sync(0xffffffffu, pend);
    const bool force = __any_sync(0xffffffffu, stuck);
    if (force || __popc(waiting) >= THRESH)
    {
        if (pend)
        {
            if (T.tgroup.y == 0u)
            {
                --dc;
                const uint32_t* e = dq + dc * 3 * kBlock;
                T.tgroup.x = e[0];
                T.tgroup.y = e[kBlock];
                T.tvalid = e[2 * kBlock];
            }
            const bool found = curvePhase ? trav_prim<2, ANY, STATS, false>(T, S.segs, rayMask, ray, hit, st)
                                          : trav_prim<1, ANY, STATS, PF>(T, S.tris, rayMask, ray, hit, st);
            if (found)
            {
                anyHit = true;
                dc = 0;
                T.tgroup.y = 0u;
                return false;
            }
        }
    }
    return active && (haveNodes || T.tgroup.y != 0u || dc > 0);
}

#ifndef SB_EXTEND_MIN_BLOCKS
#define SB_EXTEND_MIN_BLOCKS 8
#endif
// how the child-box bytes become floats (wide_node_hits<XU>), measured per kernel (profiles/r02_ab_experiments.txt,
// r2t..r2x): fp16-pair unpack with the z planes of two children per half on the XU pipe for the closest-hit loop and the
// one-ray-per-thread kernels (SB_SIMPLE_XU); byte permutes with the z pair of every child on the XU pipe for the any-hit
// loop (the fp16 form costs it 3-7 % on the 2 M-triangle scene)
#ifndef SB_EXTEND_XU
#define SB_EXTEND_XU (kXuHalfUnpack | 0x24)
#endif
#ifndef SB_SHADOW_XU
#define SB_SHADOW_XU (16 * 4 + 4)
#endif
#ifndef SB_EXTEND_PREFETCH
#define SB_EXTEND_PREFETCH 1 // closest-hit rays prefetch the next pending triangle of a leaf group (trav_prim PF)
#endif
template <bool STATS, bool CURVES>
__global__ void __launch_bounds__(kBlock, SB_EXTEND_MIN_BLOCKS) k_extend(FrameParams P, SceneDev S, Queues Q, uint32_t depth)
{
    const uint32_t n = Q.counts[count_path(depth)];
    uint32_t* head = &Q.counts[kHeadExtendBase + depth];
    const bool haveTris = S.numTriNodes != 0u, haveSegs = CURVES && S.numSegNodes != 0u; // CURVES = false: no curve code at all
    TravStats st = { 0, 0, 0, 0 };
    bool active = false;
    WarpFetch wf;
    uint32_t slot = 0;
    int phase = 0;
    Ray ray;
    RayPrep rp;
    HitRec hit;
    Traversal T;
    TravStack K;
    constexpr int kDeferDepth = CURVES ? SB_DEFER_CURVES : SB_DEFER_TRIS, kDeferThresh = CURVES ? SB_DEFER_CURVES_THRESH : SB_DEFER_TRIS_THRESH;
    __shared__ uint32_t s_defer[kDeferDepth > 0 ? kDeferDepth * 3 * kBlock : 1];
    uint32_t* dq = s_defer + threadIdx.x;
    int dc = 0;
#if SB_SMEM_STACK
    static_assert(kTravBlock == kBlock, "shared-memory traversal stack stride");
    __shared__ uint2 s_stack[SB_SMEM_STACK * kBlock];
    T.sstack = s_stack + threadIdx.x;
#endif
#if SB_PERM_SMEM && SB_FIXED_BITS
    __shared__ uint32_t s_perm[512];
    for (uint32_t i = threadIdx.x; i < 512u; i += kBlock)
        s_perm[i] = reinterpret_cast<const uint32_t*>(g_octantPerm.v)[i];
    __syncthreads();
    const uint32_t permBase = uint32_t(__cvta_generic_to_shared(s_perm));
#else
    const uint32_t permBase = 0u;
#endif
    for (;;)
    {
        const uint32_t got = fetch_slots(wf, head, n, !active, Q.rayO[0], Q.rayD[0], nullptr);
        if (got != 0xffffffffu)
        {
            slot = got;
            const float4 ro = Q.rayO[0][slot], rd = Q.rayD[0][slot];
            ray.o = mk3(ro);
            ray.d = mk3(rd);
            ray.tmin = P.materialTmin;
            ray.tmax = 1e16f;
            rp = prepare_ray(ray.d, permBase);
            hit.t = hit.u = hit.v = 0.0f;
            hit.prim = hit.inst = hit.kind = 0u;
            hit.gid = 0xffffffffu;
            trav_init(T);
            phase = haveTris ? 0 : (haveSegs ? 1 : 2);
            active = true;
        }
        if (__ballot_sync(0xffffffffu, active) == 0u)
            break;
        for (;;)
        {
            bool more = false, anyHit = false;
            if (kDeferDepth > 0)
                more = trav_iter_defer<false, STATS, CURVES, (SB_EXTEND_PREFETCH != 0), SB_EXTEND_XU, kDeferDepth, kDeferThresh>(active, phase, S, T, K, dq, dc, kRayMaskPrimary, ray, rp, hit, anyHit, &st);
            if (active)
            {
                // step shape: node visit + one primitive test per iteration.  Measured on the 2 M / 10 M-triangle and hair
                // scenes against one unit per iteration (best for the any-hit kernel) and node + all its primitives
                // (+7 % here; best for the one-ray-per-thread kernels, which have no refill ballots).
                if (kDeferDepth > 0)
                {
                }
                else if (phase == 0)
                    more = trav_step<1, false, STATS, (SB_SMEM_STACK > 0), (SB_EXTEND_PREFETCH != 0), SB_EXTEND_XU>(T, K, S.triNodes, S.tris, kRayMaskPrimary, ray, rp, hit, anyHit, &st);
                else if (CURVES && phase == 1)
                    more = trav_step<2, false, STATS, (SB_SMEM_STACK > 0), false, SB_EXTEND_XU>(T, K, S.segNodes, S.segs, kRayMaskPrimary, ray, rp, hit, anyHit, &st);
                if (!more)
                {
                    if (phase == 0 && haveSegs)
                    {
                        phase = 1;
                        trav_init(T);
                    }
                    else
                    {
                        if (CURVES && hit.kind == 2u)
                        {
                            const SegInfo si = S.segInfo[hit.prim];
                            hit.inst = si.inst;
                            hit.u = span_to_segment_u(si.span, hit.u);
                        }
                        Q.hitA[slot] = mk4(hit.t, hit.u, hit.v, u2f(hit.kind == 1u ? hit.gid : hit.prim));
                        Q.hitB[slot] = hit.inst | (hit.kind << 30);
                        active = false;
                    }
                }
            }
            const int busy = __popc(__ballot_sync(0xffffffffu, active));
            if (busy == 0 || (!(wf.exhausted && wf.next == wf.end) && busy < kRefill))
                break;
        }
    }
    if (STATS)
        flush_stats(Q.stats, st, false);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&Q.stats->radianceRays, (unsigned long long)n);
}

#ifndef SB_SHADOW_MIN_BLOCKS
#define SB_SHADOW_MIN_BLOCKS 8
#endif
// PF: next-triangle prefetch (trav_prim): chosen per scene by the launcher -- on when the triangle records exceed the L2
template <bool STATS, bool CURVES, bool PF = false>
__global__ void __launch_bounds__(kBlock, SB_SHADOW_MIN_BLOCKS) k_shadow(SceneDev S, Queues Q, uint32_t depth)
{
    const uint32_t n = Q.counts[count_shadow(depth)];
    uint32_t* head = &Q.counts[kHeadShadowBase + depth];
    const bool haveTris = S.numTriNodes != 0u, haveSegs = CURVES && S.numSegNodes != 0u; // CURVES = false: no curve code at all
    TravStats st = { 0, 0, 0, 0 };
    bool active = false;
    WarpFetch wf;
    uint32_t slot = 0;
    int phase = 0;
    Ray ray;
    RayPrep rp;
    HitRec hit;
    Traversal T;
    TravStack K;
    constexpr int kDeferDepth = CURVES ? SB_DEFER_CURVES : SB_DEFER_TRIS_SHADOW, kDeferThresh = CURVES ? SB_DEFER_CURVES_THRESH : SB_DEFER_TRIS_THRESH;
    __shared__ uint32_t s_defer[kDeferDepth > 0 ? kDeferDepth * 3 * kBlock : 1];
    uint32_t* dq = s_defer + threadIdx.x;
    int dc = 0;
#if SB_SMEM_STACK
    static_assert(kTravBlock == kBlock, "shared-memory traversal stack stride");
    __shared__ uint2 s_stack[SB_SMEM_STACK * kBlock];
    T.sstack = s_stack + threadIdx.x;
#endif
#if SB_PERM_SMEM && SB_FIXED_BITS
    __shared__ uint32_t s_perm[512];
    for (uint32_t i = threadIdx.x; i < 512u; i += kBlock)
        s_perm[i] = reinterpret_cast<const uint32_t*>(g_octantPerm.v)[i];
    __syncthreads();
    const uint32_t permBase = uint32_t(__cvta_generic_to_shared(s_perm));
#else
    const uint32_t permBase = 0u;
#endif
    for (;;)
    {
        const uint32_t got = fetch_slots(wf, head, n, !active, Q.shO, Q.shD, Q.shC);
        if (got != 0xffffffffu)
        {
            slot = got;
            const float4 so = Q.shO[slot], sd = Q.shD[slot];
            ray.o = mk3(so);
            ray.tmin = so.w;
            ray.d = mk3(sd);
            ray.tmax = sd.w;
            rp = prepare_ray(ray.d, permBase);
            hit.kind = 0u;
            hit.gid = 0xffffffffu;
            trav_init(T);
            phase = haveTris ? 0 : (haveSegs ? 1 : 2);
            active = true;
        }
        if (__ballot_sync(0xffffffffu, active) == 0u)
            break;
        for (;;)
        {
            bool more = false, occluded = false;
            if (kDeferDepth > 0)
                more = trav_iter_defer<true, STATS, CURVES, PF, SB_SHADOW_XU, kDeferDepth, kDeferThresh>(active, phase, S, T, K, dq, dc, kRayMaskShadow, ray, rp, hit, occluded, &st);
            if (active)
            {
                // step shape: ONE unit (a primitive test if one is pending, else a node visit) per iteration: measured
                // 1.5x faster for any-hit rays than node + primitive (profiles/r01_b_*)
                if (kDeferDepth > 0)
                {
                }
                else if (phase == 0)
                    more = trav_step_unit<1, true, STATS, (SB_SMEM_STACK > 0), PF, SB_SHADOW_XU>(T, K, S.triNodes, S.tris, kRayMaskShadow, ray, rp, hit, occluded, &st);
                else if (CURVES && phase == 1)
                    more = trav_step_unit<2, true, STATS, (SB_SMEM_STACK > 0), false, SB_SHADOW_XU>(T, K, S.segNodes, S.segs, kRayMaskShadow, ray, rp, hit, occluded, &st);
                if (!more)
                {
                    if (!occluded && phase == 0 && haveSegs)
                    {
                        phase = 1;
                        trav_init(T);
                    }
                    else
                    {
                        if (!occluded)
                        {
                            const float4 sc = Q.shC[slot];
                            const uint32_t pathId = f2u(sc.w);
                            const float4 L = Q.Lacc[pathId];
                            Q.Lacc[pathId] = mk4(L.x + sc.x, L.y + sc.y, L.z + sc.z, L.w);
                        }
                        active = false;
                    }
                }
            }
            const int busy = __popc(__ballot_sync(0xffffffffu, active));
            if (busy == 0 || (!(wf.exhausted && wf.next == wf.end) && busy < kRefill))
                break;
        }
    }
    if (STATS)
        flush_stats(Q.stats, st, true);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&Q.stats->shadowRays, (unsigned long long)n);
}

// shade_bounce sink of k_shade: radiance goes straight to Lacc, the shadow ray is parked in shared memory
struct StagedSink
{
    const Queues& Q;
    float4 *o, *d, *c;
    bool shadow;
    __device__ __forceinline__ void radiance_changed(const PathState& ps) const
    {
        Q.Lacc[ps.pathId] = ps.L;
    }
    __device__ __forceinline__ void shadow_ray(const PathState& ps, const float4& so, const float4& sd, const float3& contrib)
    {
        *o = so;
        *d = sd;
        *c = mk4(contrib, u2f(ps.pathId));
        shadow = true;
    }
};

#ifndef SB_SHADE_MIN_BLOCKS
#define SB_SHADE_MIN_BLOCKS 8 // measured: 64 registers + a few L1 spills beat 111 registers at 25 % occupancy (latency-bound kernel)
#endif
#ifndef SB_SHADE_HAIR_MIN_BLOCKS
#define SB_SHADE_HAIR_MIN_BLOCKS 4 // the fibre BSDF wants ~112 registers (no spills); measured +1.4 % on the C4 step
#endif
template <bool CURVES, bool PREVIEW, bool RECT_UNIFORM, bool HAIR = false, bool TEX = false>
__global__ void __launch_bounds__(kBlock, (HAIR ? SB_SHADE_HAIR_MIN_BLOCKS : SB_SHADE_MIN_BLOCKS)) k_shade(FrameParams P, SceneDev S, Queues Q, uint32_t depth)
{
    // 12 KB of byte-sliced Sobol tables per CTA (L2-resident source; 6 x 128-bit loads per thread)
    __shared__ __align__(16) uint32_t s_tab[kSobolTabWords];
    for (uint32_t i = threadIdx.x; i < kSobolTabWords / 4; i += blockDim.x)
        reinterpret_cast<uint4*>(s_tab)[i] = __ldg(reinterpret_cast<const uint4*>(Q.sobolTab) + i);
    // 4 KB table of the 1024 possible unpacked normal components (exact same expression as the scalar path)
    __shared__ float s_unpack[kUnpackLutSize];
    for (uint32_t i = threadIdx.x; i < kUnpackLutSize; i += blockDim.x)
        s_unpack[i] = unpack_component(i);
    __syncthreads();
    // shadow rays wait in shared memory (not in registers: the kernel is register-bound) until the block
    // allocates its queue slots at the end of the iteration.  (Measured alternative: one 64-bit atomic per warp
    // reserving both queues at once -- the two cursors share a word, see count_path -- with the next records
    // prefetched while it is in flight: no barriers, but 2 % slower than the block-aggregated form.  Prefetching
    // the next five queue records ahead of the barriers in THIS form: +5 % time, the 17 extra live registers spill.)
    __shared__ float4 s_shO[kBlock], s_shD[kBlock], s_shC[kBlock];
    __shared__ BlockAlloc s_alloc;
    const uint32_t n = Q.counts[count_path(depth)];
    uint32_t buf = 0;
    for (uint32_t base = blockIdx.x * kBlock; base < n; base += gridDim.x * kBlock, buf ^= 1u)
    {
        const uint32_t slot = base + threadIdx.x;
        StagedSink sink = { Q, &s_shO[threadIdx.x], &s_shD[threadIdx.x], &s_shC[threadIdx.x], false };
        PathState ps;
        bool next = false;
        if (slot < n)
        {
            const uint32_t hb = Q.hitB[slot];
            if ((hb >> 30) != 0u) // else __miss__ms: the path ends
            {
                const float4 ro = Q.rayO[0][slot], rd = Q.rayD[0][slot], th = Q.thr[0][slot];
                const float4 ha = Q.hitA[slot];
                ps.o = mk3(ro);
                ps.d = mk3(rd);
                ps.lastBsdfPdf = rd.w;
                ps.throughput = mk3(th);
                ps.flags = f2u(th.w);
                ps.pathId = f2u(ro.w);
                ps.L = Q.Lacc[ps.pathId];
                next = shade_bounce<CURVES, PREVIEW, RECT_UNIFORM, HAIR, TEX>(P, S, ps, ha, hb, depth, s_tab, s_unpack, sink);
            }
        }
        uint32_t sslot, nslot;
        block_alloc2(s_alloc, buf, &Q.counts[count_shadow(depth)], &Q.counts[count_path(depth + 1u)], sink.shadow, next, sslot, nslot);
        if (sink.shadow)
        {
            Q.shO[sslot] = s_shO[threadIdx.x];
            Q.shD[sslot] = s_shD[threadIdx.x];
            Q.shC[sslot] = s_shC[threadIdx.x];
        }
        if (next)
        {
            Q.rayO[1][nslot] = mk4(ps.o, u2f(ps.pathId));
            Q.rayD[1][nslot] = mk4(ps.d, ps.lastBsdfPdf);
            Q.thr[1][nslot] = mk4(ps.throughput, u2f(ps.flags));
        }
    }
}

// ---- single-kernel path tracer for small scenes (opt-in: SB_CFG_FUSED_SMALL) -------------------------------
// An experiment kept for A/B runs.  When the BVH is a handful of nodes (the Cornell box of the headline
// benchmark: 3 nodes, 36 triangles) nothing of the scene comes from DRAM and the wavefront form streams ~300 B of
// queue records per ray through HBM, so one would expect a register-resident megakernel to win.  Here the path
// state stays in registers for the whole path and a lane whose path has ended takes the next (pixel, sample)
// immediately (path regeneration).  Path ids are handed out in runs of kFusedGrab per warp (one L2 atomic per
// run).  Per path the arithmetic is that of the wavefront kernels, operation for operation: bit-identical images.
// MEASURED on B200, C2 at 64 spp: 42.5 ms (80 registers, 6 blocks/SM; 46.7 ms at 128 registers, 44.4 ms at 64)
// against 30.9 ms for the wavefront form.  Regeneration keeps the lanes occupied but not useful: only ~45 % of
// the bounces cast a shadow ray and ~10 % of the rays miss, so most lanes idle through the shadow traversal and
// the tail of shading, which the wavefront queues compact away; and with 6 warps per scheduler the chains of
// dependent L1 loads are not hidden.  The wavefront form stays the default for every scene size.
#ifndef SB_FUSED_MIN_BLOCKS
#define SB_FUSED_MIN_BLOCKS 6
#endif
constexpr uint32_t kFusedGrab = 256;
template <bool STATS>
__global__ void __launch_bounds__(kBlock, SB_FUSED_MIN_BLOCKS) k_path_fused(FrameParams P, SceneDev S, Queues Q)
{
    __shared__ __align__(16) uint32_t s_tab[kSobolTabWords];
    for (uint32_t i = threadIdx.x; i < kSobolTabWords / 4; i += blockDim.x)
        reinterpret_cast<uint4*>(s_tab)[i] = __ldg(reinterpret_cast<const uint4*>(Q.sobolTab) + i);
    __shared__ float s_unpack[kUnpackLutSize];
    for (uint32_t i = threadIdx.x; i < kUnpackLutSize; i += blockDim.x)
        s_unpack[i] = unpack_component(i);
    __syncthreads();
    const uint32_t nPaths = P.nPixPadded * P.chunk;
    uint32_t* head = &Q.counts[kHeadFused];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned laneLt = (1u << lane) - 1u;
    uint32_t runNext = 0, runEnd = 0; // warp-uniform: the run of path ids this warp still owns
    bool exhausted = false; // warp-uniform
    bool alive = false;
    uint32_t depth = 0;
    PathState ps;
    uint32_t nRadiance = 0, nShadow = 0;
    TravStats stE = { 0, 0, 0, 0 }, stS = { 0, 0, 0, 0 };
    for (;;)
    {
        // ---- regeneration: hand the next path ids to the idle lanes
        unsigned want = __ballot_sync(0xffffffffu, !alive);
        while (want != 0u)
        {
            if (runNext == runEnd)
            {
                if (exhausted)
                    break;
                uint32_t base = 0;
                if (lane == 0u)
                    base = atomicAdd(head, kFusedGrab);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= nPaths)
                {
                    exhausted = true;
                    break;
                }
                runNext = base;
                runEnd = min(base + kFusedGrab, nPaths);
            }
            const uint32_t avail = runEnd - runNext;
            const uint32_t rank = __popc(want & laneLt);
            if (!alive && rank < avail)
            {
                const uint32_t pathId = runNext + rank;
                alive = raygen_state(P, pathId, ps);
                depth = 0u;
                if (!alive)
                    Q.Lacc[pathId] = mk4(0.0f, 0.0f, 0.0f, 0.0f); // padding pixel of an edge tile
            }
            runNext += min(avail, uint32_t(__popc(want)));
            want = __ballot_sync(0xffffffffu, !alive);
        }
        if (__ballot_sync(0xffffffffu, alive) == 0u)
            break;
        // ---- one bounce for every lane that holds a path
        if (alive)
        {
            ++nRadiance;
            const bool next = path_bounce<STATS>(P, S, ps, depth, s_tab, s_unpack, nShadow, &stE, &stS);
            if (next)
            {
                ++depth;
            }
            else
            {
                Q.Lacc[ps.pathId] = ps.L;
                alive = false;
            }
        }
    }
    if (STATS)
    {
        flush_stats(Q.stats, stE, false);
        flush_stats(Q.stats, stS, true);
    }
    for (int off = 16; off > 0; off >>= 1)
    {
        nRadiance += __shfl_xor_sync(0xffffffffu, nRadiance, off);
        nShadow += __shfl_xor_sync(0xffffffffu, nShadow, off);
    }
    if (lane == 0u)
    {
        atomicAdd(&Q.stats->radianceRays, (unsigned long long)nRadiance);
        atomicAdd(&Q.stats->shadowRays, (unsigned long long)nShadow);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&Q.stats->paths, (unsigned long long)(P.width) * P.height * P.chunk);
}

// ---- one-ray-per-thread variants -------------------------------------------------------------------------
// Coherent or very short traversals (primary rays; scenes whose whole BVH is a handful of nodes) finish within a
// few steps of each other: there the dynamic-fetch machinery only adds ballots and an L2 atomic per refill.
// Grid sizes in CTAs per SM.  The grid-stride kernels assign slots to CTAs statically, so with exactly one resident
// wave (8 CTAs/SM) the slowest CTA sets the kernel time; more, smaller CTA workloads let the block scheduler balance
// them.  Measured on C2: one-ray-per-thread traversal kernels 16 -> 256 CTAs/SM: -8 %; shade 8 -> 128: -4 % (each
// shade CTA first fills 16 KB of shared-memory tables, so more than that costs again); 1024+: launch overhead shows.
// The persistent kernels balance themselves through the dynamic ray fetch and stay at one wave.
#ifndef SB_FUSED_PRIMARY
#define SB_FUSED_PRIMARY 1
#endif
#ifndef SB_ACC_GRID
#define SB_ACC_GRID 32
#endif
#ifndef SB_RAYGEN_GRID
#define SB_RAYGEN_GRID 8
#endif
#ifndef SB_SHADE_GRID
#define SB_SHADE_GRID 128
#endif
#ifndef SB_SIMPLE_GRID
#define SB_SIMPLE_GRID 256
#endif
#ifndef SB_SIMPLE_MIN_BLOCKS
#define SB_SIMPLE_MIN_BLOCKS 8
#endif
template <bool STATS, bool CURVES>
__global__ void __launch_bounds__(kBlock, SB_SIMPLE_MIN_BLOCKS) k_extend_simple(FrameParams P, SceneDev S, Queues Q, uint32_t depth)
{
    const uint32_t n = Q.counts[count_path(depth)];
    TravStats st = { 0, 0, 0, 0 };
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        extend_one<STATS, CURVES>(P, S, Q, 0, i, &st);
    if (STATS)
        flush_stats(Q.stats, st, false);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&Q.stats->radianceRays, (unsigned long long)n);
}

template <bool STATS, bool CURVES>
__global__ void __launch_bounds__(kBlock, SB_SIMPLE_MIN_BLOCKS) k_shadow_simple(SceneDev S, Queues Q, uint32_t depth)
{
    const uint32_t n = Q.counts[count_shadow(depth)];
    TravStats st = { 0, 0, 0, 0 };
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        shadow_one<STATS, CURVES>(S, Q, i, &st);
    if (STATS)
        flush_stats(Q.stats, st, true);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&Q.stats->shadowRays, (unsigned long long)n);
}

__global__ void __launch_bounds__(256) k_accumulate(FrameParams P, Queues Q, AccumTargets A, uint32_t mode, uint32_t subframe, uint32_t launchSamples,
                                                    uint32_t batchFlags)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P.nPixPadded; p += gridDim.x * blockDim.x)
        accumulate_pixel(P, Q, A.S, A.direct, A.aovD, A.aovS, mode, subframe, p, launchSamples, batchFlags, A.scrD, A.scrS);
}

// format: SB_FORMAT_* of the output buffer
__global__ void __launch_bounds__(256) k_resolve(const float4* S, void* out, uint32_t npix, uint32_t n, float3 e, uint32_t tonemapper, float gamma,
                                                 uint32_t format)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x)
    {
        const float4 c = resolve_pixel(S[i], n, e, tonemapper, gamma);
        if (format == SB_FORMAT_FLOAT4)
        {
            reinterpret_cast<float4*>(out)[i] = c;
        }
        else if (format == SB_FORMAT_FLOAT3)
        {
            float* o = reinterpret_cast<float*>(out) + 3 * size_t(i);
            o[0] = c.x;
            o[1] = c.y;
            o[2] = c.z;
        }
        else
        {
            uchar4 q;
            q.x = (unsigned char)(saturate(c.x) * 255.0f + 0.5f);
            q.y = (unsigned char)(saturate(c.y) * 255.0f + 0.5f);
            q.z = (unsigned char)(saturate(c.z) * 255.0f + 0.5f);
            q.w = 255;
            reinterpret_cast<uchar4*>(out)[i] = q;
        }
    }
}

// debug / no-accumulation output: the launch result goes to the output buffer, through the same post-process
// as accumulated frames (the reference tonemaps params.image whatever wrote it, OptixRender.cpp:1045-1049)
__global__ void __launch_bounds__(256) k_copy_image(const float4* src, void* out, uint32_t npix, uint32_t format, float3 e, uint32_t tonemapper,
                                                    float gamma)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x)
    {
        const float4 c = postprocess_pixel(mk3(src[i]), e, tonemapper, gamma);
        if (format == SB_FORMAT_FLOAT4)
        {
            reinterpret_cast<float4*>(out)[i] = c;
        }
        else if (format == SB_FORMAT_FLOAT3)
        {
            float* o = reinterpret_cast<float*>(out) + 3 * size_t(i);
            o[0] = c.x;
            o[1] = c.y;
            o[2] = c.z;
        }
        else
        {
            uchar4 q;
            q.x = (unsigned char)(saturate(c.x) * 255.0f + 0.5f);
            q.y = (unsigned char)(saturate(c.y) * 255.0f + 0.5f);
            q.z = (unsigned char)(saturate(c.z) * 255.0f + 0.5f);
            q.w = 255;
            reinterpret_cast<uchar4*>(out)[i] = q;
        }
    }
}

// ---- launchers -------------------------------------------------------------------------------------

cudaEvent_t StageTimer::get()
{
    if (!pool.empty())
    {
        cudaEvent_t e = pool.back();
        pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    SB_CUDA_CHECK(cudaEventCreate(&e));
    return e;
}
void StageTimer::collect()
{
    for (const Pending& p : pending)
    {
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess)
        {
            ms[p.stage] += t;
            launches[p.stage] += 1;
        }
        else
        {
            cudaGetLastError();
        }
        pool.push_back(p.a);
        pool.push_back(p.b);
    }
    pending.clear();
}
void StageTimer::reset()
{
    collect();
    for (int i = 0; i < kNumStages; ++i)
    {
        ms[i] = 0.0;
        launches[i] = 0;
    }
}
void StageTimer::release()
{
    collect();
    for (cudaEvent_t e : pool)
        cudaEventDestroy(e);
    pool.clear();
}

// counts the launch and, when stage timers are on, brackets it with events
struct ScopedStage
{
    const LaunchCfg& cfg;
    StageTimer::Pending p;
    bool on;
    ScopedStage(const LaunchCfg& c, int stage) : cfg(c), on(c.timer != nullptr)
    {
        if (cfg.launchCount)
            ++*cfg.launchCount;
        if (on)
        {
            p.stage = stage;
            p.a = cfg.timer->get();
            p.b = cfg.timer->get();
            cudaEventRecord(p.a, cfg.stream);
        }
    }
    ~ScopedStage()
    {
        if (on)
        {
            cudaEventRecord(p.b, cfg.stream);
            cfg.timer->pending.push_back(p);
        }
    }
};

static inline unsigned grid_for(const LaunchCfg& cfg, int blocksPerSm)
{
    return unsigned(cfg.numSms * blocksPerSm);
}

// run-time flags -> compile-time template arguments: calls f(std::bool_constant<a>, <b>, <c>)
template <class F>
static void launch_shade_variant(bool a, bool b, bool c, F f)
{
    auto with_c = [&](auto ta, auto tb) {
        if (c)
            f(ta, tb, std::true_type());
        else
            f(ta, tb, std::false_type());
    };
    auto with_b = [&](auto ta) {
        if (b)
            with_c(ta, std::true_type());
        else
            with_c(ta, std::false_type());
    };
    if (a)
        with_b(std::true_type());
    else
        with_b(std::false_type());
}

// Closest-hit stage of bounce `depth` over path queue 0 of `Q` (the caller has already swapped the ping-pong pointers).
// Camera rays (depth 0) are coherent and a scene of a few nodes is traversed in a few steps: one ray per thread;
// everything else runs the persistent dynamic-fetch kernel.  Scenes without curves run kernels compiled without the
// curve phase (the traversal loop is very sensitive to its size and shape).
void launch_extend_stage(const LaunchCfg& cfg, const FrameParams& P, const SceneDev& S, const Queues& Q, uint32_t depth, bool stats)
{
    ScopedStage sc(cfg, kStageExtend);
    cudaStream_t st = cfg.stream;
    const bool tiny = (S.numTriNodes + S.numSegNodes) <= kTinyBvhNodes;
    const bool curves = S.numSegNodes != 0u;
    const bool persistent = !tiny && depth > 0;
    if (persistent)
    {
        if (curves)
        {
            if (stats)
                k_extend<true, true><<<grid_for(cfg, SB_EXTEND_MIN_BLOCKS), kBlock, 0, st>>>(P, S, Q, depth);
            else
                k_extend<false, true><<<grid_for(cfg, SB_EXTEND_MIN_BLOCKS), kBlock, 0, st>>>(P, S, Q, depth);
        }
        else
        {
            if (stats)
                k_extend<true, false><<<grid_for(cfg, SB_EXTEND_MIN_BLOCKS), kBlock, 0, st>>>(P, S, Q, depth);
            else
                k_extend<false, false><<<grid_for(cfg, SB_EXTEND_MIN_BLOCKS), kBlock, 0, st>>>(P, S, Q, depth);
        }
    }
    else
    {
        if (curves)
        {
            if (stats)
                k_extend_simple<true, true><<<grid_for(cfg, SB_SIMPLE_GRID), kBlock, 0, st>>>(P, S, Q, depth);
            else
                k_extend_simple<false, true><<<grid_for(cfg, SB_SIMPLE_GRID), kBlock, 0, st>>>(P, S, Q, depth);
        }
        else
        {
            if (stats)
                k_extend_simple<true, false><<<grid_for(cfg, SB_SIMPLE_GRID), kBlock, 0, st>>>(P, S, Q, depth);
            else
                k_extend_simple<false, false><<<grid_for(cfg, SB_SIMPLE_GRID), kBlock, 0, st>>>(P, S, Q, depth);
        }
    }
    SB_CUDA_CHECK(cudaGetLastError());
}

// Any-hit stage over the shadow queue of bounce `depth`.
void launch_shadow_stage(const LaunchCfg& cfg, const SceneDev& S, const Queues& Q, uint32_t depth, bool stats)
{
    ScopedStage sc(cfg, kStageShadow);
    cudaStream_t st = cfg.stream;
    const bool tiny = (S.numTriNodes + S.numSegNodes) <= kTinyBvhNodes;
    const bool curves = S.numSegNodes != 0u;
    if (!tiny)
    {
        // any-hit rays prefetch their next pending triangle only where the records do not fit the L2 (they often stop
        // before it: a loss of 13 % on the 2 M-triangle scene, a gain of 6 % on the 10 M-triangle one)
        const bool prefetch = size_t(S.numTris) * sizeof(TriRec) > kPrefetchTriBytes;
        if (curves)
        {
            if (stats)
                k_shadow<true, true><<<grid_for(cfg, SB_SHADOW_MIN_BLOCKS), kBlock, 0, st>>>(S, Q, depth);
            else
                k_shadow<false, true><<<grid_for(cfg, SB_SHADOW_MIN_BLOCKS), kBlock, 0, st>>>(S, Q, depth);
        }
        else if (prefetch)
        {
            if (stats)
                k_shadow<true, false, true><<<grid_for(cfg, SB_SHADOW_MIN_BLOCKS), kBlock, 0, st>>>(S, Q, depth);
            else
                k_shadow<false, false, true><<<grid_for(cfg, SB_SHADOW_MIN_BLOCKS), kBlock, 0, st>>>(S, Q, depth);
        }
        else
        {
            if (stats)
                k_shadow<true, false><<<grid_for(cfg, SB_SHADOW_MIN_BLOCKS), kBlock, 0, st>>>(S, Q, depth);
            else
                k_shadow<false, false><<<grid_for(cfg, SB_SHADOW_MIN_BLOCKS), kBlock, 0, st>>>(S, Q, depth);
        }
    }
    else
    {
        if (curves)
        {
            if (stats)
                k_shadow_simple<true, true><<<grid_for(cfg, SB_SIMPLE_GRID), kBlock, 0, st>>>(S, Q, depth);
            else
                k_shadow_simple<false, true><<<grid_for(cfg, SB_SIMPLE_GRID), kBlock, 0, st>>>(S, Q, depth);
        }
        else
        {
            if (stats)
                k_shadow_simple<true, false><<<grid_for(cfg, SB_SIMPLE_GRID), kBlock, 0, st>>>(S, Q, depth);
            else
                k_shadow_simple<false, false><<<grid_for(cfg, SB_SIMPLE_GRID), kBlock, 0, st>>>(S, Q, depth);
        }
    }
    SB_CUDA_CHECK(cudaGetLastError());
}

void launch_wavefront_batch(const LaunchCfg& cfg, const FrameParams& P, const SceneDev& S, const Queues& Qbase, bool stats)
{
    const Queues& Q = Qbase;
    cudaStream_t st = cfg.stream;
    const bool tiny = (S.numTriNodes + S.numSegNodes) <= kTinyBvhNodes;
    const bool curves = S.numSegNodes != 0u;
    SB_CUDA_CHECK(cudaMemsetAsync(Q.counts, 0, sizeof(uint32_t) * kNumCounts, st));
    if (tiny && cfg.fusedSmall)
    {
        ScopedStage sc(cfg, kStagePathFused);
        if (stats)
            k_path_fused<true><<<grid_for(cfg, SB_FUSED_MIN_BLOCKS), kBlock, 0, st>>>(P, S, Q);
        else
            k_path_fused<false><<<grid_for(cfg, SB_FUSED_MIN_BLOCKS), kBlock, 0, st>>>(P, S, Q);
        SB_CUDA_CHECK(cudaGetLastError());
        return;
    }
    const bool fusedPrimary = SB_FUSED_PRIMARY && P.nPixPadded == P.width * P.height; // no padding pixels
    if (fusedPrimary)
    {
        ScopedStage sc(cfg, kStageExtend);
        if (curves)
        {
            if (stats)
                k_primary<true, true><<<grid_for(cfg, SB_SIMPLE_GRID), kBlock, 0, st>>>(P, S, Q);
            else
                k_primary<false, true><<<grid_for(cfg, SB_SIMPLE_GRID), kBlock, 0, st>>>(P, S, Q);
        }
        else
        {
            if (stats)
                k_primary<true, false><<<grid_for(cfg, SB_SIMPLE_GRID), kBlock, 0, st>>>(P, S, Q);
            else
                k_primary<false, false><<<grid_for(cfg, SB_SIMPLE_GRID), kBlock, 0, st>>>(P, S, Q);
        }
    }
    else
    {
        ScopedStage sc(cfg, kStageRaygen);
        k_raygen<<<grid_for(cfg, SB_RAYGEN_GRID), kBlock, 0, st>>>(P, Q);
    }
    for (uint32_t depth = 0; depth < P.maxDepth; ++depth)
    {
        // the kernels read path queue 0 and write path queue 1: swap the ping-pong pointers per bounce
        Queues Q = Qbase;
        if (depth & 1u)
        {
            std::swap(Q.rayO[0], Q.rayO[1]);
            std::swap(Q.rayD[0], Q.rayD[1]);
            std::swap(Q.thr[0], Q.thr[1]);
        }
        if (depth > 0 || !fusedPrimary)
            launch_extend_stage(cfg, P, S, Q, depth, stats);
        {
            ScopedStage sc(cfg, kStageShade);
            // the variant without the code paths this scene cannot take
            const bool preview = S.anyPreviewMaterial, rectUniform = S.onlyRectLights && P.rectMethod == 0u;
            if (S.numTextures != 0u)
            {
                // textured scenes run the general variant (every material model, texture lookups)
                if (rectUniform)
                    k_shade<true, true, true, true, true><<<grid_for(cfg, SB_SHADE_GRID), kBlock, 0, st>>>(P, S, Q, depth);
                else
                    k_shade<true, true, false, true, true><<<grid_for(cfg, SB_SHADE_GRID), kBlock, 0, st>>>(P, S, Q, depth);
            }
            else if (S.anyHairMaterial)
            {
                // scenes with a hair material run the general variant (curves, every material model)
                if (rectUniform)
                    k_shade<true, true, true, true><<<grid_for(cfg, SB_SHADE_GRID), kBlock, 0, st>>>(P, S, Q, depth);
                else
                    k_shade<true, true, false, true><<<grid_for(cfg, SB_SHADE_GRID), kBlock, 0, st>>>(P, S, Q, depth);
            }
            else
                launch_shade_variant(curves, preview, rectUniform, [&](auto c, auto p, auto r) {
                    k_shade<decltype(c)::value, decltype(p)::value, decltype(r)::value><<<grid_for(cfg, SB_SHADE_GRID), kBlock, 0, st>>>(P, S, Q, depth);
                });
        }
        if (P.debug == 1u)
            break; // debug normals: only the first hit is shaded (OptixRender.cu:151-152)
        launch_shadow_stage(cfg, S, Q, depth, stats);
    }
    SB_CUDA_CHECK(cudaGetLastError());
}

void launch_accumulate(const LaunchCfg& cfg, const FrameParams& P, const Queues& Q, const AccumTargets& A, uint32_t mode, uint32_t subframe,
                       uint32_t launchSamples, uint32_t batchFlags)
{
    ScopedStage sc(cfg, kStageAccumulate);
    k_accumulate<<<grid_for(cfg, SB_ACC_GRID), 256, 0, cfg.stream>>>(P, Q, A, mode, subframe, launchSamples, batchFlags);
    SB_CUDA_CHECK(cudaGetLastError());
}

void launch_resolve(const LaunchCfg& cfg, const float4* S, void* out, uint32_t npix, uint32_t n, const float exposure[3], uint32_t tonemapper,
                    float gamma, uint32_t format)
{
    ScopedStage sc(cfg, kStageResolve);
    k_resolve<<<grid_for(cfg, 4), 256, 0, cfg.stream>>>(S, out, npix, n, make_float3(exposure[0], exposure[1], exposure[2]), tonemapper, gamma, format);
    SB_CUDA_CHECK(cudaGetLastError());
}

void launch_copy_image(const LaunchCfg& cfg, const float4* src, void* out, uint32_t npix, uint32_t format, const float exposure[3],
                       uint32_t tonemapper, float gamma)
{
    ScopedStage sc(cfg, kStageResolve);
    k_copy_image<<<grid_for(cfg, 4), 256, 0, cfg.stream>>>(src, out, npix, format, make_float3(exposure[0], exposure[1], exposure[2]), tonemapper, gamma);
    SB_CUDA_CHECK(cudaGetLastError());
}

// ---- test hooks ------------------------------------------------------------------------------------

__global__ void k_test_sampler(uint32_t n, const uint32_t* x, const uint32_t* y, const uint32_t* sample, const uint32_t* maxs, const uint32_t* depth,
                               const uint32_t* dim, float* out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t sidx = sampler_index(x[i], y[i], sample[i], maxs[i]);
        const float a = sampler_rnd(sidx, depth[i], dim[i]);
        // cross-check the fused five-value path the integrator uses against the generic one
        const Sample5 s5 = sampler_sample5(sidx, depth[i]);
        const float b = s5.v[(dim[i] + depth[i] * 10u) % 5u];
        out[i] = (a == b) ? a : -1.0f;
    }
}

__global__ void k_test_light_sample(uint32_t n, const sb_light* lights, const float* hp, const float* u, uint32_t method, float* out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
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

__global__ void k_test_trace(SceneDev S, uint32_t n, const float* rays, uint32_t mode, sb_hit* hits)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
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
                traverse_bvh<1, false, false>(S.triNodes, S.tris, kRayMaskPrimary, ray, rp, hit, &st);
            if (S.numSegNodes)
            {
                if (traverse_bvh<2, false, false>(S.segNodes, S.segs, kRayMaskPrimary, ray, rp, hit, &st))
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
                occ = traverse_bvh<1, true, false>(S.triNodes, S.tris, kRayMaskShadow, ray, rp, hit, &st);
            if (!occ && S.numSegNodes)
                occ = traverse_bvh<2, true, false>(S.segNodes, S.segs, kRayMaskShadow, ray, rp, hit, &st);
            h.kind = occ ? 1u : 0u;
        }
        hits[i] = h;
    }
}

// ---- sb_test_trace modes 2 / 3: the caller's rays through the REAL queues and the production stage launchers --------
__global__ void k_test_fill_extend(Queues Q, uint32_t n, const float* rays, uint32_t depth)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float* r = rays + 8 * size_t(i);
        Q.rayO[0][i] = mk4(r[0], r[1], r[2], u2f(i));
        Q.rayD[0][i] = mk4(r[4], r[5], r[6], 0.0f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        Q.counts[count_path(depth)] = n;
}
__global__ void k_test_read_hits(SceneDev S, Queues Q, const uint32_t* instTriFirst, uint32_t n, sb_hit* hits)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float4 ha = Q.hitA[i];
        const uint32_t hb = Q.hitB[i];
        sb_hit h;
        h.t = ha.x;
        h.u = ha.y;
        h.v = ha.z;
        h.kind = hb >> 30;
        h.instance = hb & 0x0fffffffu;
        h.prim = 0u;
        if (h.kind == 1u)
            h.prim = f2u(ha.w) - instTriFirst[h.instance]; // global triangle id -> index inside the instance's mesh
        else if (h.kind == 2u)
            h.prim = S.segInfo[f2u(ha.w)].prim;
        hits[i] = h;
    }
}
__global__ void k_test_fill_shadow(Queues Q, uint32_t n, const float* rays, uint32_t depth)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float* r = rays + 8 * size_t(i);
        Q.shO[i] = mk4(r[0], r[1], r[2], r[3]);
        Q.shD[i] = mk4(r[4], r[5], r[6], r[7]);
        Q.shC[i] = mk4(1.0f, 0.0f, 0.0f, u2f(i));
        Q.Lacc[i] = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        Q.counts[count_shadow(depth)] = n;
}
__global__ void k_test_read_occlusion(Queues Q, uint32_t n, sb_hit* hits)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        sb_hit h = { 0, 0, 0, 0, 0, 0 };
        h.kind = Q.Lacc[i].x == 0.0f ? 1u : 0u; // the unoccluded ones received their contribution of 1
        hits[i] = h;
    }
}

void launch_test_trace_production(const LaunchCfg& cfg, const FrameParams& P, const SceneDev& S, const Queues& Q, const uint32_t* instTriFirst,
                                  uint32_t n, const float* rays, uint32_t mode, sb_hit* hits, bool stats)
{
    const uint32_t depth = 1u; // any bounce after the camera rays: the dispatch of secondary rays
    SB_CUDA_CHECK(cudaMemsetAsync(Q.counts, 0, sizeof(uint32_t) * kNumCounts, cfg.stream));
    if (mode == 2u)
    {
        k_test_fill_extend<<<grid_for(cfg, 2), 128, 0, cfg.stream>>>(Q, n, rays, depth);
        launch_extend_stage(cfg, P, S, Q, depth, stats);
        k_test_read_hits<<<grid_for(cfg, 2), 128, 0, cfg.stream>>>(S, Q, instTriFirst, n, hits);
    }
    else
    {
        k_test_fill_shadow<<<grid_for(cfg, 2), 128, 0, cfg.stream>>>(Q, n, rays, depth);
        launch_shadow_stage(cfg, S, Q, depth, stats);
        k_test_read_occlusion<<<grid_for(cfg, 2), 128, 0, cfg.stream>>>(Q, n, hits);
    }
    SB_CUDA_CHECK(cudaGetLastError());
}

__global__ void k_test_offset_ray(uint32_t n, const float* p, const float* nrm, float* out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float3 r = offset_ray(mk3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), mk3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
        out[3 * i] = r.x;
        out[3 * i + 1] = r.y;
        out[3 * i + 2] = r.z;
    }
}
void launch_test_offset_ray(const LaunchCfg& cfg, uint32_t n, const float* p, const float* nrm, float* out)
{
    k_test_offset_ray<<<grid_for(cfg, 2), 128, 0, cfg.stream>>>(n, p, nrm, out);
    SB_CUDA_CHECK(cudaGetLastError());
}

// sb_test_texture: n lookups of texture index1 (1-based) at (u, v) -> rgba
__global__ void k_test_texture(SceneDev S, uint32_t index1, uint32_t n, const float* uv, float* out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float4 c = sample_texture(S, index1, uv[2 * i], uv[2 * i + 1]);
        out[4 * i] = c.x;
        out[4 * i + 1] = c.y;
        out[4 * i + 2] = c.z;
        out[4 * i + 3] = c.w;
    }
}
void launch_test_texture(const LaunchCfg& cfg, const SceneDev& S, uint32_t index1, uint32_t n, const float* uv, float* out)
{
    k_test_texture<<<grid_for(cfg, 2), 128, 0, cfg.stream>>>(S, index1, n, uv, out);
    SB_CUDA_CHECK(cudaGetLastError());
}

// sb_test_bsdf: the BSDF protocol (sample + evaluate) on caller inputs; 19 floats in, 15 floats out per item
__global__ void k_test_bsdf(sb_material m, uint32_t n, const float* in, float* out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float* a = in + 19 * size_t(i);
        float* o = out + 15 * size_t(i);
        const float3 N = mk3(a[0], a[1], a[2]), NG = mk3(a[3], a[4], a[5]), T = mk3(a[6], a[7], a[8]), K1 = mk3(a[9], a[10], a[11]);
        const float3 base = mk3(m.base_color[0], m.base_color[1], m.base_color[2]);
        const BsdfSample s = bsdf_sample<true, true>(m, base, N, NG, T, K1, mk4(a[12], a[13], a[14], a[15]));
        o[0] = s.k2.x;
        o[1] = s.k2.y;
        o[2] = s.k2.z;
        o[3] = s.bsdf_over_pdf.x;
        o[4] = s.bsdf_over_pdf.y;
        o[5] = s.bsdf_over_pdf.z;
        o[6] = s.pdf;
        o[7] = float(s.event);
        const BsdfEval e = bsdf_evaluate<true, true>(m, base, N, NG, T, K1, mk3(a[16], a[17], a[18]));
        o[8] = e.diffuse.x;
        o[9] = e.diffuse.y;
        o[10] = e.diffuse.z;
        o[11] = e.glossy.x;
        o[12] = e.glossy.y;
        o[13] = e.glossy.z;
        o[14] = e.pdf;
    }
}
void launch_test_bsdf(const LaunchCfg& cfg, const sb_material& m, uint32_t n, const float* in, float* out)
{
    k_test_bsdf<<<grid_for(cfg, 4), 128, 0, cfg.stream>>>(m, n, in, out);
    SB_CUDA_CHECK(cudaGetLastError());
}

void launch_test_sampler(const LaunchCfg& cfg, uint32_t n, const uint32_t* x, const uint32_t* y, const uint32_t* sample, const uint32_t* maxs,
                         const uint32_t* depth, const uint32_t* dim, float* out)
{
    k_test_sampler<<<grid_for(cfg, 2), 128, 0, cfg.stream>>>(n, x, y, sample, maxs, depth, dim, out);
    SB_CUDA_CHECK(cudaGetLastError());
}
void launch_test_light_sample(const LaunchCfg& cfg, uint32_t n, const sb_light* lights, const float* hp, const float* u, uint32_t method, float* out)
{
    k_test_light_sample<<<grid_for(cfg, 2), 128, 0, cfg.stream>>>(n, lights, hp, u, method, out);
    SB_CUDA_CHECK(cudaGetLastError());
}
void launch_test_trace(const LaunchCfg& cfg, const SceneDev& S, uint32_t n, const float* rays, uint32_t mode, sb_hit* hits)
{
    k_test_trace<<<grid_for(cfg, 2), 128, 0, cfg.stream>>>(S, n, rays, mode, hits);
    SB_CUDA_CHECK(cudaGetLastError());
}

} // namespace sb
