// Software traversal of the compressed 8-wide BVH: closest-hit and any-hit queries over world-space
// triangles (Moller-Trumbore on precomputed edges) and round cubic B-spline curve segments.
// Replaces optixTrace (OptixRender.cu:120-129: mask 255, FLAG_NONE, no culling; closest_hit.cu:185-197:
// mask RAY_MASK_SHADOW=3, TERMINATE_ON_FIRST_HIT) -- closed driver code in the reference.
//
// Determinism / parity rule shared with the CPU oracle: among hits with the same t the primitive with
// the lower global id (instance-major, primitive-minor order of the scene arrays) wins, so the reported
// hit does not depend on the shape of the acceleration structure; triangles beat curves on exact ties.
#pragma once
#include "bvh.cuh"
#include "curve.cuh"

#ifndef SB_SMEM_STACK
#define SB_SMEM_STACK 8
#endif
#ifndef SB_F32X2
#define SB_F32X2 1
#endif
#ifndef SB_SIMPLE_WW
#define SB_SIMPLE_WW 1
#endif

namespace sb
{

#if defined(__CUDA_ARCH__)
#define SB_LDG4(p) __ldg(reinterpret_cast<const uint4*>(p))
#define SB_LDGF4(p) __ldg(reinterpret_cast<const float4*>(p))
#else
#define SB_LDG4(p) (*reinterpret_cast<const uint4*>(p))
#define SB_LDGF4(p) (*reinterpret_cast<const float4*>(p))
#endif

// GEOMETRY_MASK_*, OptixRenderParams.h:9-17
constexpr uint32_t kMaskTriangle = 1u, kMaskCurve = 2u, kMaskLight = 4u;
constexpr uint32_t kRayMaskPrimary = 255u, kRayMaskShadow = 3u;

struct TriRec // 48 B
{
    float4 v0; // xyz, w = bits(primitive index inside its mesh)
    float4 e1; // v1 - v0, w = bits(instance | visibilityMask << 28)
    float4 e2; // v2 - v0, w = bits(global triangle id)
};
struct SegRec // 64 B: power-basis coefficients of one span of a curve segment in world space (curve.cuh)
{
    float4 q[4];
};
struct SegInfo
{
    uint32_t prim; // optixGetPrimitiveIndex: index into the curve prim's segment list
    uint32_t inst;
    uint32_t firstPoint; // global index of the first control point
    uint32_t span; // k | K << 16: this record covers u in [k/K, (k+1)/K] of the segment
};
// segment parameter of a hit at span-local parameter s
SB_HD float span_to_segment_u(uint32_t span, float s)
{
    const float K = float(span >> 16);
    return (float(span & 0xffffu) + s) * (1.0f / K);
}

struct Ray
{
    float3 o;
    float tmin;
    float3 d;
    float tmax;
};
struct HitRec
{
    float t, u, v;
    uint32_t prim, inst, kind; // kind 0 miss, 1 triangle, 2 curve
    uint32_t gid;
};
struct TravStats
{
    uint32_t nodes, tris, segs, overflow;
};

constexpr int kStackSize = 32;

SB_HD uint32_t byte_of(uint32_t w, int j)
{
    return (w >> (8 * j)) & 0xffu;
}
// float(byte j of w), exactly, without an int->float conversion: I2F runs on the XU pipe (16 lanes/clk/SM) and
// was 59 % busy in the first ncu capture of this kernel.  One byte-permute builds the bit pattern of
// 2^23 + b, one FADD removes the 2^23 -- both steps are exact, so the result equals float(b) bit for bit.
// (Measured alternative: fold the offset into the slab coefficients, t = (1 + b 2^-15) * (2^15 a) + (b0 - 2^15 a),
// one FFMA per plane and no FADD.  The folded constant is only accurate to 2^-9 of a quantisation cell, so the
// planes need that much slack; rays leaving an axis-aligned wall then enter the flat boxes of its co-planar
// neighbours: +38 % triangle tests on the Cornell box (-3 % Mrays/s), for -4 % traversal time on the 2 M-triangle
// scene.  Kept exact.  Also measured: uncompressed float child boxes (224-byte nodes) for scenes of <= 64 nodes --
// no conversions at all, but the 9 extra vector loads and the pointer selects cost as much ALU work as the 48 PRMTs
// they remove: +2 % on the Cornell box.)
SB_HD float byte_to_float(uint32_t w, int j)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + uint32_t(j))) - 8388608.0f;
#else
    return float(byte_of(w, j));
#endif
}

#ifndef SB_MAGIC_CONST
#define SB_MAGIC_CONST 1
#endif
#ifndef SB_HALF_UNPACK
#define SB_HALF_UNPACK 1 // two bytes per byte-permute through an fp16 pair (byte_pair_to_float); needs SB_FIXED_BITS
#endif
// XU template argument of wide_node_hits: how the quantised child planes become floats, chosen per kernel.
//   without kXuHalfUnpack: 16 * pairs + n -- PRMT + FADD2 per byte, except the pairs `pairs` (1: near xy, 2: far xy,
//     4: z) of the first n children of each half, which go through I2F.U8 on the XU pipe
//   with kXuHalfUnpack: fp16-pair unpack (byte_pair_to_float), except plane word i (0..2 near x y z, 3..5 far x y z) of
//     children 0,1 (bit i) / children 2,3 (bit 8 + i) of each half, which go through I2F.U8
constexpr int kXuHalfUnpack = 0x10000;
#if SB_MAGIC_CONST && defined(__CUDACC__)
static __constant__ uint32_t c_byteMagic = 0x4B000000u;
static __constant__ uint32_t c_halfMagic = 0x64646464u; // SB_HALF_UNPACK
// FHADD reads its fp32 addend from a vector register only; as a literal or a constant-bank value ptxas re-materialises
// it per use (40 moves per node visit).  Loaded from (mutable) global memory once per ray it stays in one register.
static __device__ float g_halfBias = -1024.0f;
#endif
#if defined(__CUDA_ARCH__)
// Packed fp32 arithmetic of sm_100 (FADD2 / FFMA2: two IEEE fp32 operations per issued instruction, each
// component rounded exactly like the scalar instruction).  The traversal kernels are bound by instruction issue,
// and the slab tests come in natural pairs (x|y near, x|y far, z near|far).
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// bit pattern of 2^23 + byte j of w (see byte_to_float).  PRMT takes ONE immediate: with the 2^23 pattern as a literal,
// ptxas spends it on that and re-materialises the four byte selectors in registers -- 65 moves per node visit in the
// SASS of the persistent kernels.  Read from constant memory the pattern is a c[3][..] operand and the selector is the
// immediate (SB_MAGIC_CONST=1).
__device__ __forceinline__ float byte_magic(uint32_t w, int j)
{
#if SB_MAGIC_CONST
    return __uint_as_float(__byte_perm(w, c_byteMagic, 0x7540u + uint32_t(j)));
#else
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + uint32_t(j)));
#endif
}
#endif

// SB_FIXED_BITS: how the hit mask of a node visit is assembled.
//   0: Ylitie et al. 2017 -- per child, (unary triangle count) << (meta index ^ octant): five ALU instructions each.
//   1: every child slot s owns FIXED bits of the mask -- 3s..3s+2 if it is a leaf (its triangles, unary), 24+s if it is
//      an inner node -- so a hit costs one predicated OR with an immediate; the node carries the valid-bit word
//      (n1.z) that strips the bits a slot does not own.  The octant permutation of the inner byte (traversal order:
//      bit 24 + (s ^ octinv)) is one lookup in a 2 KB table, and a pending triangle's index is recovered from the
//      valid word with one population count: primBase + popc(valid below the bit).
#ifndef SB_FIXED_BITS
#define SB_FIXED_BITS 1
#endif
#if SB_FIXED_BITS
struct OctantPermLut
{
    uint8_t v[8 * 256]; // v[o * 256 + x]: bit s of x moved to bit s ^ o
    constexpr OctantPermLut() : v{}
    {
        for (int o = 0; o < 8; ++o)
            for (int x = 0; x < 256; ++x)
            {
                int y = 0;
                for (int b = 0; b < 8; ++b)
                    if (x & (1 << b))
                        y |= 1 << (b ^ o);
                v[o * 256 + x] = uint8_t(y);
            }
    }
};
#if defined(__CUDACC__)
static __device__ const OctantPermLut g_octantPerm = OctantPermLut();
#endif
// inner: hit inner children by slot (8 bits); octinv8 = octinv << 8 (+ the shared-memory address of a copy of the
// table when SMEM: the persistent kernels stage one per CTA, SB_PERM_SMEM)
#ifndef SB_PERM_SMEM
#define SB_PERM_SMEM 1
#endif
template <bool SMEM = false>
SB_HD uint32_t permute_inner_hits(uint32_t inner, uint32_t octinv8)
{
#if defined(__CUDA_ARCH__)
    if (SMEM)
    {
        uint32_t v;
        asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(octinv8 + inner));
        return v;
    }
    return __ldg(&g_octantPerm.v[octinv8 | inner]);
#else
    const uint32_t o = octinv8 >> 8;
    uint32_t y = 0;
    for (uint32_t b = 0; b < 8u; ++b)
        if (inner & (1u << b))
            y |= 1u << (b ^ o);
    return y;
#endif
}
#endif

#if defined(__CUDA_ARCH__)
// SB_HALF_UNPACK: bytes j and j+1 of w -> two floats through ONE byte-permute: the permute builds the fp16 pair
// (1024 + b_j, 1024 + b_j+1) (0x64 in the high bytes: 1024 has an ulp of 1), and the mixed-precision add of sm_100
// (add.f32.f16 -> FHADD, fma pipe, reads either half of a register) removes the 1024 in fp32.  Exact like byte_magic,
// with half the ALU-pipe instructions.
__device__ __forceinline__ void byte_pair_to_float(uint32_t w, int j, float bias, float& q0, float& q1)
{
    const uint32_t h = __byte_perm(w, c_halfMagic, j == 0 ? 0x4140u : 0x4342u);
    asm("{ .reg .f16 l, h; mov.b32 {l, h}, %2; add.rn.f32.f16 %0, l, %3; add.rn.f32.f16 %1, h, %3; }"
        : "=f"(q0), "=f"(q1)
        : "r"(h), "f"(bias));
}
#endif

// Ray/child-box tests of one wide node -> 32-bit hit mask: bits 24..31 inner children in traversal
// priority (highest first), bits 0..23 leaf primitives relative to primBase.
template <int XU = 0>
SB_HD uint32_t wide_node_hits(const uint4& n0, const uint4& n1, const uint4& n2, const uint4& n3, const uint4& n4, const float3& o,
                              const float3& idir, uint32_t octinv4, bool negx, bool negy, bool negz, float tmin, float tmax, float halfBias = -1024.0f)
{
    const float3 p = mk3(u2f(n0.x), u2f(n0.y), u2f(n0.z));
    // child plane distance t = q * (2^e * idir) + (p - o) * idir
    const float ax = u2f((n0.w & 0xffu) << 23) * idir.x;
    const float ay = u2f(((n0.w >> 8) & 0xffu) << 23) * idir.y;
    const float az = u2f(((n0.w >> 16) & 0xffu) << 23) * idir.z;
    const float bx = (p.x - o.x) * idir.x;
    const float by = (p.y - o.y) * idir.y;
    const float bz = (p.z - o.z) * idir.z;
    uint32_t hitmask = 0;
#if defined(__CUDA_ARCH__) && SB_F32X2
    const unsigned long long Axy = f2_pack(ax, ay), Bxy = f2_pack(bx, by), Azz = f2_pack(az, az), Bzz = f2_pack(bz, bz);
    const unsigned long long kMagic = f2_pack(-8388608.0f, -8388608.0f);
#endif
#pragma unroll
    for (int half = 0; half < 2; ++half)
    {
#if !SB_FIXED_BITS
        const uint32_t meta4 = half ? n1.w : n1.z;
        const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t innerMask4 = (isInner4 >> 4) * 0xffu;
        const uint32_t bitIndex4 = (meta4 ^ (octinv4 & innerMask4)) & 0x1f1f1f1fu;
        const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
#endif
        const uint32_t qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z, qloz = half ? n3.y : n3.x;
        const uint32_t qhix = half ? n3.w : n3.z, qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
        const uint32_t nearx = negx ? qhix : qlox, farx = negx ? qlox : qhix;
        const uint32_t neary = negy ? qhiy : qloy, fary = negy ? qloy : qhiy;
        const uint32_t nearz = negz ? qhiz : qloz, farz = negz ? qloz : qhiz;
#if defined(__CUDA_ARCH__) && SB_F32X2 && SB_HALF_UNPACK && SB_FIXED_BITS
        if (XU & kXuHalfUnpack)
        {
            // two children per step.  XU: bit i (children 0,1) / bit 8 + i (children 2,3) of each half sends plane word i
            // (0..2 near x y z, 3..5 far x y z) through I2F.U8 on the XU pipe instead
#pragma unroll
            for (int jp = 0; jp < 4; jp += 2)
            {
                float t0x[2], t0y[2], t0z[2], t1x[2], t1y[2], t1z[2];
#define SB_SLAB2(idx, w, a, b, r)                                                                                          \
        {                                                                                                                  \
            float q0, q1;                                                                                                  \
            if ((XU >> (idx + 4 * jp)) & 1)                                                                                \
            {                                                                                                              \
                q0 = float(byte_of(w, jp));                                                                                \
                q1 = float(byte_of(w, jp + 1));                                                                            \
            }                                                                                                              \
            else                                                                                                           \
                byte_pair_to_float(w, jp, halfBias, q0, q1);                                                               \
            f2_unpack(f2_fma(f2_pack(q0, q1), f2_pack(a, a), f2_pack(b, b)), r[0], r[1]);                                  \
        }
                SB_SLAB2(0, nearx, ax, bx, t0x)
                SB_SLAB2(1, neary, ay, by, t0y)
                SB_SLAB2(2, nearz, az, bz, t0z)
                SB_SLAB2(3, farx, ax, bx, t1x)
                SB_SLAB2(4, fary, ay, by, t1y)
                SB_SLAB2(5, farz, az, bz, t1z)
#undef SB_SLAB2
#pragma unroll
                for (int k = 0; k < 2; ++k)
                {
                    const float cmin = fmaxf(fmaxf(t0x[k], t0y[k]), fmaxf(t0z[k], tmin));
                    const float cmax = fminf(fminf(t1x[k], t1y[k]), fminf(t1z[k], tmax)) * 1.0000004f;
                    if (cmin <= cmax)
                        hitmask |= (7u << (3 * (4 * half + jp + k))) | (1u << (24 + 4 * half + jp + k));
                }
            }
            continue;
        }
#endif
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
#if defined(__CUDA_ARCH__) && SB_F32X2
            float t0x, t0y, t0z, t1x, t1y, t1z;
            // XU = 16 * pairs + n: the pairs selected by `pairs` (1: near xy, 2: far xy, 4: z) of the first n children of
            // each half convert on the otherwise idle XU pipe -- I2F.U8 with a byte selector, one instruction instead of
            // PRMT + half an FADD2 (the ALU pipe is the busiest one in this loop).  Both conversions are exact.
#define SB_SLAB_PAIR(bit, wa, wb, A, B, ra, rb)                                                                           \
    if (((XU >> 4) & bit) != 0 && j < (XU & 15))                                                                           \
        f2_unpack(f2_fma(f2_pack(float(byte_of(wa, j)), float(byte_of(wb, j))), A, B), ra, rb);                            \
    else                                                                                                                   \
        f2_unpack(f2_fma(f2_add(f2_pack(byte_magic(wa, j), byte_magic(wb, j)), kMagic), A, B), ra, rb);
            SB_SLAB_PAIR(1, nearx, neary, Axy, Bxy, t0x, t0y)
            SB_SLAB_PAIR(2, farx, fary, Axy, Bxy, t1x, t1y)
            SB_SLAB_PAIR(4, nearz, farz, Azz, Bzz, t0z, t1z)
#undef SB_SLAB_PAIR
#else
            const float t0x = fmaf(byte_to_float(nearx, j), ax, bx);
            const float t0y = fmaf(byte_to_float(neary, j), ay, by);
            const float t0z = fmaf(byte_to_float(nearz, j), az, bz);
            const float t1x = fmaf(byte_to_float(farx, j), ax, bx);
            const float t1y = fmaf(byte_to_float(fary, j), ay, by);
            const float t1z = fmaf(byte_to_float(farz, j), az, bz);
#endif
            const float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
            // conservative far plane: the 1 + 2*gamma(3) factor of Ize 2013 (the CPU oracle's slab test uses it
            // too).  It must scale the RESULT: pre-scaled plane coefficients lose it to cancellation.
            const float cmax = fminf(fminf(t1x, t1y), fminf(t1z, tmax)) * 1.0000004f;
#if SB_FIXED_BITS
            if (cmin <= cmax)
                hitmask |= (7u << (3 * (4 * half + j))) | (1u << (24 + 4 * half + j));
#else
            if (cmin <= cmax)
                hitmask |= byte_of(childBits4, j) << byte_of(bitIndex4, j);
#endif
        }
    }
#if SB_FIXED_BITS
    return hitmask & n1.z; // inner bits by SLOT (not yet in traversal order)
#else
    return hitmask;
#endif
}

// Moller-Trumbore on (v0, e1, e2); expression tree identical to the oracle's (fma dot/cross).
SB_HD bool intersect_tri(const float3& v0, const float3& e1, const float3& e2, const float3& o, const float3& d, float tmin, float tmax,
                         float& t, float& u, float& v)
{
    const float3 p = cross_fma(d, e2);
    const float det = dot_fma(e1, p);
    if (det == 0.0f)
        return false;
    const float inv = 1.0f / det;
    const float3 tv = o - v0;
    u = dot_fma(tv, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f))
        return false;
    const float3 q = cross_fma(tv, e1);
    v = dot_fma(d, q) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f))
        return false;
    t = dot_fma(e2, q) * inv;
    return t > tmin && t <= tmax;
}

struct RayPrep
{
    float3 idir;
    uint32_t octinv4;
    uint32_t octinv;
    bool negx, negy, negz;
    float halfBias; // -1024 in a register the compiler cannot re-materialise (byte_pair_to_float)
};
SB_HD RayPrep prepare_ray(const float3& d, uint32_t permLutBase = 0u)
{
    RayPrep r;
    r.negx = d.x < 0.0f;
    r.negy = d.y < 0.0f;
    r.negz = d.z < 0.0f;
    const uint32_t oct = (r.negx ? 1u : 0u) | (r.negy ? 2u : 0u) | (r.negz ? 4u : 0u);
    r.octinv = 7u - oct;
#if defined(__CUDA_ARCH__)
    r.halfBias = g_halfBias;
#else
    r.halfBias = -1024.0f;
#endif
#if SB_FIXED_BITS
    r.octinv4 = permLutBase + (r.octinv << 8); // row of the permutation table
#else
    r.octinv4 = r.octinv * 0x01010101u;
#endif
    const float eps = 1e-20f;
    r.idir.x = 1.0f / (fabsf(d.x) > eps ? d.x : copysignf(eps, d.x));
    r.idir.y = 1.0f / (fabsf(d.y) > eps ? d.y : copysignf(eps, d.y));
    r.idir.z = 1.0f / (fabsf(d.z) > eps ? d.z : copysignf(eps, d.z));
    return r;
}

// Traversal state of one ray in one BVH.  A "node group" is (child base index, hit bits | inner mask), a
// "primitive group" is (primitive base index, hit bits); the stack holds postponed node groups only.
// SB_SMEM_STACK > 0: the first SB_SMEM_STACK stack levels of the persistent kernels live in shared memory
// (strided by the block size, bank-conflict free), deeper ones in local memory.
constexpr int kTravBlock = 128; // threads per block of the kernels that use the shared-memory stack
// The postponed node groups live in their own array (TravStack), NOT inside Traversal: with a dynamically indexed array
// member nvcc keeps the whole object addressable and writes ngroup / tgroup / sp back to local memory after every
// update (4 STL per node visit in the SASS of the persistent kernels, 31 M local stores per 7 M-ray launch in ncu).
struct TravStack
{
    uint2 e[kStackSize];
};
struct Traversal
{
    uint2 ngroup, tgroup;
    uint32_t tvalid; // SB_FIXED_BITS: valid-bit word of the node the pending primitives belong to
    int sp;
#if defined(__CUDACC__)
    uint2* sstack; // this thread's column of the block's shared-memory stack (SSTACK traversals only)
#endif
};
#define SB_TSTACK(T, K) (K).e
SB_HD void trav_init(Traversal& T)
{
    T.ngroup.x = 0u;
    T.ngroup.y = 0x80000000u; // the root: one inner "child" at base 0
    T.tgroup.x = 0u;
    T.tgroup.y = 0u;
    T.sp = 0;
}

// Traversal is driven in two half-steps so that a warp of incoherent rays interleaves node visits and
// primitive tests lane by lane, while coherent rays stay in lock-step:
//   trav_node : if the lane has no primitives pending, visit the next node (returns false when nothing is left)
//   trav_prim : if the lane has primitives pending, test exactly one
// KIND 1: triangles (prims = TriRec), KIND 2: curve segments (prims = SegRec).  ANY: shadow rays.
#ifndef SB_SIMPLE_XU
#define SB_SIMPLE_XU (kXuHalfUnpack | 0x24) // conversion mix of the byte conversions (wide_node_hits) in the one-ray-per-thread traversals
#endif
#ifndef SB_PF_ONE
#define SB_PF_ONE 1 // one prefetch (the middle of the 48-byte record) instead of two (first and last word): C3 extend -2.6 %, C5 equal
#endif
#ifndef SB_SIMPLE_PREFETCH
#define SB_SIMPLE_PREFETCH 0 // next-triangle prefetch in the one-ray-per-thread closest-hit traversal (camera rays)
#endif
template <bool STATS, bool SSTACK = false, int XU = 0>
SB_HD bool trav_node(Traversal& T, TravStack& K, const WideNode* __restrict__ nodes, const Ray& ray, const RayPrep& rp, TravStats* st)
{
    if (T.ngroup.y <= 0x00ffffffu)
    {
        if (T.sp == 0)
            return false;
        --T.sp;
#if defined(__CUDA_ARCH__) && SB_SMEM_STACK
        if (SSTACK)
            T.ngroup = (T.sp < SB_SMEM_STACK) ? T.sstack[T.sp * kTravBlock] : SB_TSTACK(T, K)[T.sp - SB_SMEM_STACK];
        else
#endif
            T.ngroup = SB_TSTACK(T, K)[T.sp];
    }
    const uint32_t hits = T.ngroup.y;
    const uint32_t bit = bfind32(hits);
    T.ngroup.y &= ~(1u << bit);
    if (T.ngroup.y > 0x00ffffffu)
    {
        if (T.sp < kStackSize)
        {
#if defined(__CUDA_ARCH__) && SB_SMEM_STACK
            if (SSTACK)
            {
                if (T.sp < SB_SMEM_STACK)
                    T.sstack[T.sp * kTravBlock] = T.ngroup;
                else
                    SB_TSTACK(T, K)[T.sp - SB_SMEM_STACK] = T.ngroup;
            }
            else
#endif
                SB_TSTACK(T, K)[T.sp] = T.ngroup;
            ++T.sp;
        }
        else if (STATS)
            st->overflow++;
    }
    const uint32_t slot = (bit - 24u) ^ (rp.octinv & 7u);
    const uint32_t rel = popc32(hits & ~(0xffffffffu << slot) & 0xffu);
    const uint32_t ni = T.ngroup.x + rel;
    const WideNode* np = nodes + ni;
    uint4 n0, n1, n2, n3, n4;
#if defined(__CUDA_ARCH__) && SB_NODE96
    uint4 unused;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(n0.x), "=r"(n0.y), "=r"(n0.z), "=r"(n0.w), "=r"(n1.x), "=r"(n1.y), "=r"(n1.z), "=r"(n1.w)
                 : "l"(&np->n0));
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(n2.x), "=r"(n2.y), "=r"(n2.z), "=r"(n2.w), "=r"(n3.x), "=r"(n3.y), "=r"(n3.z), "=r"(n3.w)
                 : "l"(&np->n2));
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(n4.x), "=r"(n4.y), "=r"(n4.z), "=r"(n4.w), "=r"(unused.x), "=r"(unused.y), "=r"(unused.z), "=r"(unused.w)
                 : "l"(&np->n4));
#else
    n0 = SB_LDG4(&np->n0);
    n1 = SB_LDG4(&np->n1);
    n2 = SB_LDG4(&np->n2);
    n3 = SB_LDG4(&np->n3);
    n4 = SB_LDG4(&np->n4);
#endif
    if (STATS)
        st->nodes++;
    const uint32_t hm = wide_node_hits<XU>(n0, n1, n2, n3, n4, ray.o, rp.idir, rp.octinv4, rp.negx, rp.negy, rp.negz, ray.tmin, ray.tmax, rp.halfBias);
    T.ngroup.x = n1.x;
#if SB_FIXED_BITS
    T.ngroup.y = (permute_inner_hits<(SSTACK && SB_PERM_SMEM != 0)>(hm >> 24, rp.octinv4) << 24) | (n0.w >> 24);
    T.tvalid = n1.z;
#else
    T.ngroup.y = (hm & 0xff000000u) | (n0.w >> 24);
#endif
    T.tgroup.x = n1.y;
    T.tgroup.y = hm & 0x00ffffffu;
    return true;
}

// offset, from the node's first primitive, of the pending primitive at bit `rel` of tgroup.y
SB_HD uint32_t prim_offset(const Traversal& T, uint32_t rel)
{
#if SB_FIXED_BITS
    return popc32(T.tvalid & ~(0xffffffffu << rel));
#else
    return rel;
#endif
}

// tests one pending primitive; returns true if an any-hit query is satisfied (ANY only)
// PF: start the fetch of the lane's NEXT pending triangle while this one is tested (pays when the triangle records
// come from DRAM: measured -6 % on the 10 M-triangle scene; on the 2 M-triangle scene, which the L2 holds, closest-hit
// rays gain 2 % and any-hit rays, which often stop before the next triangle, lose 13 %)
template <int KIND, bool ANY, bool STATS, bool PF = false>
SB_HD bool trav_prim(Traversal& T, const void* __restrict__ prims, uint32_t rayMask, Ray& ray, HitRec& hit, TravStats* st)
{
    const uint32_t rel = bfind32(T.tgroup.y);
    T.tgroup.y &= ~(1u << rel);
    const uint32_t pi = T.tgroup.x + prim_offset(T, rel);
    if (KIND == 1)
    {
        const TriRec* tr = reinterpret_cast<const TriRec*>(prims) + pi;
        const float4 a = SB_LDGF4(&tr->v0), b = SB_LDGF4(&tr->e1), c = SB_LDGF4(&tr->e2);
#if defined(__CUDA_ARCH__)
        if (PF && T.tgroup.y != 0u)
        {
            const TriRec* nx = reinterpret_cast<const TriRec*>(prims) + (T.tgroup.x + prim_offset(T, bfind32(T.tgroup.y)));
#if SB_PF_ONE
            asm volatile("prefetch.global.L1 [%0];" ::"l"(&nx->e1));
#else
            asm volatile("prefetch.global.L1 [%0];" ::"l"(&nx->v0));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(&nx->e2));
#endif
        }
#endif
        if (STATS)
            st->tris++;
        const uint32_t instMask = f2u(b.w);
        if ((instMask >> 28) & rayMask)
        {
            float t, u, v;
            if (intersect_tri(mk3(a), mk3(b), mk3(c), ray.o, ray.d, ray.tmin, ray.tmax, t, u, v))
            {
                if (ANY)
                    return true;
                const uint32_t gid = f2u(c.w);
                if (t < ray.tmax || hit.kind != 1u || gid < hit.gid)
                {
                    hit.t = t;
                    hit.u = u;
                    hit.v = v;
                    hit.prim = f2u(a.w);
                    hit.inst = instMask & 0x0fffffffu;
                    hit.kind = 1u;
                    hit.gid = gid;
                    ray.tmax = t;
                }
            }
        }
    }
    else
    {
        const SegRec* sr = reinterpret_cast<const SegRec*>(prims) + pi;
#if defined(__CUDA_ARCH__)
        if (PF && T.tgroup.y != 0u)
        {
            const SegRec* nx = reinterpret_cast<const SegRec*>(prims) + (T.tgroup.x + prim_offset(T, bfind32(T.tgroup.y)));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(&nx->q[0]));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(&nx->q[2]));
        }
#endif
        float4 q[4];
        q[0] = SB_LDGF4(&sr->q[0]);
        q[1] = SB_LDGF4(&sr->q[1]);
        q[2] = SB_LDGF4(&sr->q[2]);
        q[3] = SB_LDGF4(&sr->q[3]);
        if (STATS)
            st->segs++;
        float t, u;
        if (intersect_round_cubic(q, ray.o, ray.d, ray.tmin, ray.tmax, t, u))
        {
            if (ANY)
                return true;
            hit.t = t;
            hit.u = u;
            hit.v = 0.0f;
            hit.prim = pi; // index into the SegInfo table; resolved by the caller
            hit.kind = 2u;
            hit.gid = pi;
            ray.tmax = t;
        }
    }
    return false;
}

// one full step of one lane: node half-step if idle, then one primitive if any is pending.
// Returns false when the traversal is finished; anyHit is set when an ANY query found an occluder.
template <int KIND, bool ANY, bool STATS, bool SSTACK = false, bool PF = false, int XU = 0>
SB_HD bool trav_step(Traversal& T, TravStack& K, const WideNode* __restrict__ nodes, const void* __restrict__ prims, uint32_t rayMask, Ray& ray,
                     const RayPrep& rp, HitRec& hit, bool& anyHit, TravStats* st)
{
    if (T.tgroup.y == 0u)
    {
        if (!trav_node<STATS, SSTACK, XU>(T, K, nodes, ray, rp, st))
            return false;
    }
    if (T.tgroup.y != 0u)
    {
        if (trav_prim<KIND, ANY, STATS, PF>(T, prims, rayMask, ray, hit, st))
        {
            anyHit = true;
            return false;
        }
    }
    return true;
}

// Single-unit variant of trav_step: ONE primitive test if any is pending, else ONE node visit.  Measured on the
// 2 M-triangle scene the any-hit (shadow) kernel runs 1.5x faster with this shape, the closest-hit kernel
// slightly faster with the node+primitive shape above (profiles/r01_b_*).
template <int KIND, bool ANY, bool STATS, bool SSTACK = false, bool PF = false, int XU = 0>
SB_HD bool trav_step_unit(Traversal& T, TravStack& K, const WideNode* __restrict__ nodes, const void* __restrict__ prims, uint32_t rayMask, Ray& ray,
                          const RayPrep& rp, HitRec& hit, bool& anyHit, TravStats* st)
{
    if (T.tgroup.y != 0u)
    {
        if (trav_prim<KIND, ANY, STATS, PF>(T, prims, rayMask, ray, hit, st))
        {
            anyHit = true;
            return false;
        }
        return true;
    }
    return trav_node<STATS, SSTACK, XU>(T, K, nodes, ray, rp, st);
}

// "While-while" step: ONE node visit, then ALL the primitives it queued.  The lanes of a warp meet again at every
// node test (the expensive half of the work) instead of drifting apart.
template <int KIND, bool ANY, bool STATS>
SB_HD bool trav_step_ww(Traversal& T, TravStack& K, const WideNode* __restrict__ nodes, const void* __restrict__ prims, uint32_t rayMask, Ray& ray,
                        const RayPrep& rp, HitRec& hit, bool& anyHit, TravStats* st)
{
    if (T.tgroup.y == 0u && !trav_node<STATS>(T, K, nodes, ray, rp, st))
        return false;
    while (T.tgroup.y != 0u)
    {
        if (trav_prim<KIND, ANY, STATS>(T, prims, rayMask, ray, hit, st))
        {
            anyHit = true;
            return false;
        }
    }
    return true;
}

// Whole traversal of one BVH for one ray (test hooks, host emulation).  Returns true if a hit was found
// (closest: hit updated and ray.tmax shrunk; any: first accepted hit).
template <int KIND, bool ANY, bool STATS>
SB_HD bool traverse_bvh(const WideNode* __restrict__ nodes, const void* __restrict__ prims, uint32_t rayMask, Ray& ray, const RayPrep& rp,
                        HitRec& hit, TravStats* st)
{
    Traversal T;
    TravStack K;
    trav_init(T);
    const uint32_t kindBefore = hit.kind;
    const float tBefore = ray.tmax;
    bool anyHit = false;
#if SB_SIMPLE_WW
    // "while-while": a lane drains the primitives of the node it just visited before the warp moves on, so the
    // lanes of a warp meet again at every node test (the expensive half) instead of drifting apart
    for (;;)
    {
        if (T.tgroup.y == 0u && !trav_node<STATS, false, SB_SIMPLE_XU>(T, K, nodes, ray, rp, st))
            break;
        while (T.tgroup.y != 0u)
        {
            if (trav_prim<KIND, ANY, STATS, (SB_SIMPLE_PREFETCH != 0 && !ANY && KIND == 1)>(T, prims, rayMask, ray, hit, st))
            {
                anyHit = true;
                break;
            }
        }
        if (ANY && anyHit)
            break;
    }
#else
    while (trav_step<KIND, ANY, STATS, false, false, SB_SIMPLE_XU>(T, K, nodes, prims, rayMask, ray, rp, hit, anyHit, st))
    {
    }
#endif
    if (ANY)
        return anyHit;
    return hit.kind == uint32_t(KIND) && (kindBefore != uint32_t(KIND) || ray.tmax < tBefore);
}

} // namespace sb
