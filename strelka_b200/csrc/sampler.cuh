// Device sampler: Morton-indexed, Owen-scrambled 5-D Sobol, bit-exact with the reference's
// src/render/optix/RandomSampler.h (initSampler :130-137, sobol_scramble :213-219, random<> :221-226).
// Parity is pinned by tests/test_gpu_parity.py::test_sampler_bit_exact_on_device (and tests/test_emul_vs_oracle.py on the
// host instantiation) against tests/golden/ref_vectors.json.
//
// The direction numbers are not copied from the reference table: sobol_generate() rebuilds them from
// the Joe-Kuo primitive polynomials of dimensions 1..5 on the host and the context uploads them to
// __constant__ memory (the bit loop below indexes them warp-uniformly -> constant-cache broadcast).
//
// Quirk Q2: the dimension is (Dim + 10*depth) % 5, so per bounce only FIVE distinct
// values exist (eBSDF0..3 alias ePixelX, ePixelY, eLightId, eLightPointX; eRussianRoulette aliases
// eLightPointY) and they share one scrambled index.  sample5() exploits exactly that: one index
// scramble + five Sobol evaluations per bounce instead of ten full draws -- same bits.
#pragma once
#include "hd.cuh"

namespace sb
{

enum SampleDim : uint32_t
{
    kPixelX = 0,
    kPixelY,
    kLightId,
    kLightPointX,
    kLightPointY,
    kBSDF0,
    kBSDF1,
    kBSDF2,
    kBSDF3,
    kRussianRoulette,
    kNumDims
};

// c_sobol is defined in kernels.cu (the library is compiled as ONE translation unit, unity.cu);
// h_sobol backs the host pass of the SB_HD functions and the g++ emulation build.
extern uint32_t h_sobol[5][32];
#if defined(__CUDA_ARCH__)
#define SB_SOBOL(d, b) c_sobol[d][b]
#else
#define SB_SOBOL(d, b) h_sobol[d][b]
#endif

// Joe-Kuo direction numbers, dims 1..5: (degree, coefficient, m_i)
inline void sobol_generate(uint32_t v[5][32])
{
    static const uint32_t deg[5] = { 0, 1, 2, 3, 3 };
    static const uint32_t coef[5] = { 0, 0, 1, 1, 2 };
    static const uint32_t minit[5][3] = { { 0, 0, 0 }, { 1, 0, 0 }, { 1, 3, 0 }, { 1, 3, 1 }, { 1, 1, 1 } };
    for (uint32_t i = 0; i < 32; ++i)
        v[0][i] = 1u << (31 - i);
    for (int d = 1; d < 5; ++d)
    {
        const uint32_t s = deg[d];
        for (uint32_t i = 0; i < 32; ++i)
        {
            if (i < s)
            {
                v[d][i] = minit[d][i] << (31 - i);
            }
            else
            {
                uint32_t x = v[d][i - s] ^ (v[d][i - s] >> s);
                for (uint32_t k = 1; k < s; ++k)
                    x ^= ((coef[d] >> (s - 1 - k)) & 1u) * v[d][i - k];
                v[d][i] = x;
            }
        }
    }
}

SB_HD uint32_t fmix32(uint32_t x) // hash(), RandomSampler.h:86-95
{
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}
SB_HD uint32_t seed_combine(uint32_t seed, uint32_t v) // hash_combine(), :50-53
{
    return seed ^ (v + (seed << 6) + (seed >> 2));
}
SB_HD uint32_t spread16(uint32_t x) // Part1By1, :115-123
{
    x &= 0x0000ffffu;
    x = (x ^ (x << 8)) & 0x00ff00ffu;
    x = (x ^ (x << 4)) & 0x0f0f0f0fu;
    x = (x ^ (x << 2)) & 0x33333333u;
    x = (x ^ (x << 1)) & 0x55555555u;
    return x;
}
SB_HD uint32_t morton2(uint32_t x, uint32_t y) // EncodeMorton2, :125-128
{
    return (spread16(y) << 1) + spread16(x);
}
// initSampler, :130-137 -- uint32 wrap-around is intended (quirk Q3)
SB_HD uint32_t sampler_index(uint32_t px, uint32_t py, uint32_t sampleIndex, uint32_t maxSamples)
{
    return morton2(px, py) * maxSamples + sampleIndex;
}
SB_HD uint32_t lk_permute(uint32_t v, uint32_t seed) // laine_karras_permutation, :182-190
{
    v += seed;
    v ^= v * 0x6c50b47cu;
    v ^= v * 0xb82f1e52u;
    v ^= v * 0xc7afe638u;
    v ^= v * 0x8d22f6e6u;
    return v;
}
SB_HD uint32_t owen_scramble(uint32_t v, uint32_t seed) // nested_uniform_scramble, :205-211
{
    return brev32(lk_permute(brev32(v), seed));
}
// Dimension 1 (polynomial x+1): the generator matrix is Pascal's triangle mod 2, column k = (1+S)^k e_31, so by
// Lucas' theorem output bit (31-j) is the parity of the index bits k that are supersets of j -- a
// superset-XOR (zeta) transform over the 5-bit positions: five masked shift-XORs and one bit reversal.
SB_HD uint32_t sobol_dim1(uint32_t g)
{
    g ^= (g >> 1) & 0x55555555u;
    g ^= (g >> 2) & 0x33333333u;
    g ^= (g >> 4) & 0x0f0f0f0fu;
    g ^= (g >> 8) & 0x00ff00ffu;
    g ^= (g >> 16) & 0x0000ffffu;
    return brev32(g);
}

// Byte-sliced tables for dimensions 2..4: tab[(d-2)*1024 + k*256 + b] = XOR of the direction numbers of the
// bits set in byte value b at byte position k.  4 lookups + 3 XORs replace the 32-step bit loop; the shade
// kernel keeps the 12 KB table in shared memory (random 4-byte gathers: ~3 bank-conflict wavefronts each).
constexpr uint32_t kSobolTabWords = 3u * 4u * 256u;
inline void sobol_build_tables(const uint32_t v[5][32], uint32_t* tab)
{
    for (uint32_t d = 2; d < 5; ++d)
        for (uint32_t k = 0; k < 4; ++k)
            for (uint32_t b = 0; b < 256; ++b)
            {
                uint32_t x = 0;
                for (uint32_t bit = 0; bit < 8; ++bit)
                    if ((b >> bit) & 1u)
                        x ^= v[d][8 * k + bit];
                tab[(d - 2) * 1024 + k * 256 + b] = x;
            }
}
SB_HD uint32_t sobol_tab(const uint32_t* tab, uint32_t index, uint32_t dim) // dim in 2..4
{
    const uint32_t* t = tab + (dim - 2u) * 1024u;
    return t[index & 0xffu] ^ t[256u + ((index >> 8) & 0xffu)] ^ t[512u + ((index >> 16) & 0xffu)] ^ t[768u + (index >> 24)];
}

SB_HD uint32_t sobol_u32(uint32_t index, uint32_t dim) // sobol_uint, :166-175
{
    if (dim == 0)
        return brev32(index); // dimension 0 is the identity matrix: van der Corput == bit reversal
    if (dim == 1)
        return sobol_dim1(index);
    uint32_t x = 0;
#pragma unroll
    for (int bit = 0; bit < 32; ++bit)
    {
        if ((index >> bit) & 1u)
            x ^= SB_SOBOL(dim, bit);
    }
    return x;
}
SB_HD float u32_to_unit(uint32_t r)
{
    return fminf(float(r) * 0x1p-32f, kOneMinusEps);
}

// random<Dim>(state), :221-226, for an arbitrary dimension enum value (test hook / generic use)
SB_HD float sampler_rnd(uint32_t sampleIdx, uint32_t depth, uint32_t dim, uint32_t seed = 52u)
{
    const uint32_t dimension = (dim + depth * uint32_t(kNumDims)) % 5u;
    uint32_t s = fmix32(seed + depth);
    const uint32_t index = owen_scramble(sampleIdx, s);
    const uint32_t r = owen_scramble(sobol_u32(index, dimension), seed_combine(s, dimension));
    return u32_to_unit(r);
}

// The five distinct values of one bounce: v[k] == random<k>() == random<k+5>() at this depth.
struct Sample5
{
    float v[5];
};
// tab: byte-sliced tables (sobol_build_tables) or nullptr for the plain bit loop -- same bits either way
SB_HD Sample5 sampler_sample5(uint32_t sampleIdx, uint32_t depth, const uint32_t* tab = nullptr, uint32_t seed = 52u)
{
    Sample5 r;
    const uint32_t s = fmix32(seed + depth);
    const uint32_t index = owen_scramble(sampleIdx, s);
    r.v[0] = u32_to_unit(owen_scramble(brev32(index), seed_combine(s, 0u)));
    r.v[1] = u32_to_unit(owen_scramble(sobol_dim1(index), seed_combine(s, 1u)));
#pragma unroll
    for (uint32_t d = 2; d < 5; ++d)
        r.v[d] = u32_to_unit(owen_scramble(tab ? sobol_tab(tab, index, d) : sobol_u32(index, d), seed_combine(s, d)));
    return r;
}

} // namespace sb
