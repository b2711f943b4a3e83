// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the reference sampler: Morton-indexed, Owen-scrambled 5-D Sobol
// (src/render/optix/RandomSampler.h).  Bit-exact parity with the reference is REQUIRED and is
// pinned by oracle/ref_crosscheck.cpp (which #includes the reference header) and by the frozen
// vectors in tests/golden/sampler.json.
//
// Nothing here is copied: the direction numbers are regenerated from the Joe-Kuo
// (new-joe-kuo-6.21201) primitive polynomials for dimensions 1..5, which is what the
// reference's literal table sb_matrix[5][32] (RandomSampler.h:139-164) contains.
#pragma once
#include <cstdint>
#include <cmath>

namespace orc
{

// The largest float below 1 (RandomSampler.h:6)
static constexpr float kOneMinusEps = 0x1.fffffep-1f;

// SampleDimension, RandomSampler.h:13-26
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

// SamplerState, RandomSampler.h:28-33
struct Sampler
{
    uint32_t seed;
    uint32_t sampleIdx;
    uint32_t depth;
};

// murmur3 finaliser == hash(), RandomSampler.h:86-95
inline uint32_t fmix32(uint32_t x)
{
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}

// hash_combine(), RandomSampler.h:50-53
inline uint32_t seed_combine(uint32_t seed, uint32_t v)
{
    return seed ^ (v + (seed << 6) + (seed >> 2));
}

// Part1By1 / EncodeMorton2, RandomSampler.h:115-128: x bit k -> bit 2k, y bit k -> bit 2k+1
inline uint32_t spread16(uint32_t x)
{
    x &= 0x0000ffffu;
    x = (x ^ (x << 8)) & 0x00ff00ffu;
    x = (x ^ (x << 4)) & 0x0f0f0f0fu;
    x = (x ^ (x << 2)) & 0x33333333u;
    x = (x ^ (x << 1)) & 0x55555555u;
    return x;
}
inline uint32_t morton2(uint32_t x, uint32_t y)
{
    return (spread16(y) << 1) + spread16(x);
}

// initSampler(), RandomSampler.h:130-137 (uint32 arithmetic wraps on purpose, quirk Q3)
inline Sampler init_sampler(uint32_t px, uint32_t py, uint32_t sampleIndex, uint32_t maxSamples, uint32_t seed)
{
    Sampler s;
    s.seed = seed;
    s.sampleIdx = morton2(px, py) * maxSamples + sampleIndex;
    s.depth = 0;
    return s;
}

// Sobol direction numbers for the first five Joe-Kuo dimensions: (degree s, coefficient a, m_i).
struct SobolTable
{
    uint32_t v[5][32];
    SobolTable()
    {
        static const uint32_t deg[5] = { 0, 1, 2, 3, 3 };
        static const uint32_t coef[5] = { 0, 0, 1, 1, 2 };
        static const uint32_t minit[5][3] = { { 0, 0, 0 }, { 1, 0, 0 }, { 1, 3, 0 }, { 1, 3, 1 }, { 1, 1, 1 } };
        for (int i = 0; i < 32; ++i)
        {
            v[0][i] = 1u << (31 - i); // van der Corput
        }
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
                    {
                        x ^= ((coef[d] >> (s - 1 - k)) & 1u) * v[d][i - k];
                    }
                    v[d][i] = x;
                }
            }
        }
    }
};
inline const SobolTable& sobol_table()
{
    static const SobolTable t;
    return t;
}

// sobol_uint(), RandomSampler.h:166-175
inline uint32_t sobol_u32(uint32_t index, uint32_t dim)
{
    const SobolTable& t = sobol_table();
    uint32_t x = 0;
    for (int bit = 0; bit < 32; ++bit)
    {
        if ((index >> bit) & 1u)
        {
            x ^= t.v[dim][bit];
        }
    }
    return x;
}

// laine_karras_permutation(), RandomSampler.h:182-190
inline uint32_t lk_permute(uint32_t v, uint32_t seed)
{
    v += seed;
    v ^= v * 0x6c50b47cu;
    v ^= v * 0xb82f1e52u;
    v ^= v * 0xc7afe638u;
    v ^= v * 0x8d22f6e6u;
    return v;
}

// ReverseBits(), RandomSampler.h:192-203
inline uint32_t bitrev32(uint32_t v)
{
    v = ((v & 0xaaaaaaaau) >> 1) | ((v & 0x55555555u) << 1);
    v = ((v & 0xccccccccu) >> 2) | ((v & 0x33333333u) << 2);
    v = ((v & 0xf0f0f0f0u) >> 4) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v & 0xff00ff00u) >> 8) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
}

// nested_uniform_scramble(), RandomSampler.h:205-211
inline uint32_t owen_scramble(uint32_t v, uint32_t seed)
{
    return bitrev32(lk_permute(bitrev32(v), seed));
}

// sobol_scramble(), RandomSampler.h:213-219
inline float sobol_owen(uint32_t index, uint32_t dim, uint32_t seed)
{
    seed = fmix32(seed);
    index = owen_scramble(index, seed);
    const uint32_t r = owen_scramble(sobol_u32(index, dim), seed_combine(seed, dim));
    return std::fmin(float(r) * 0x1p-32f, kOneMinusEps);
}

// random<Dim>(state), RandomSampler.h:221-226.  NB quirk Q2: (Dim + depth*10) % 5.
inline float rnd(SampleDim dim, const Sampler& s)
{
    const uint32_t dimension = (uint32_t(dim) + s.depth * uint32_t(kNumDims)) % 5u;
    return sobol_owen(s.sampleIdx, dimension, s.seed + s.depth);
}

} // namespace orc
