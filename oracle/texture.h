// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Software model of the texture lookups of the reference's textured materials: the MDL runtime's
// tex_lookup_float4_2d (texture_support_cuda.h:287-314) calls tex2D<float4> on a texture object created with wrap
// addressing, linear filtering, normalised coordinates and normalised-float reads of 8-bit RGBA texels
// (OptiXRender::loadTextureFromFile, OptixRender.cpp:1191-1268).  What the hardware computes is specified in the CUDA
// C++ Programming Guide, appendix "Texture Fetching", linear filtering of a two-dimensional texture:
//     tex(x, y) = (1-a)(1-b) T[i,j] + a(1-b) T[i+1,j] + (1-a) b T[i,j+1] + a b T[i+1,j+1]
//     xB = x - 0.5, i = floor(xB), a = frac(xB)   (x = N * frac(u) in wrap mode; likewise y, j, b)
//     a and b are stored in 9-bit fixed point format with 8 bits of fractional value
// This file states that definition in double precision with table-driven texel decoding; it shares no code with
// strelka_b200/csrc (the device samples the texture unit).  PARITY: hardware == this model is pinned on the device by
// tests/test_gpu_textures.py (sb_test_texture), to the tolerance the fixed-point weights allow.
#pragma once
#include "vec.h"
#include <cmath>
#include <cstdint>
#include <vector>

namespace orc
{

struct Texture
{
    std::vector<uint8_t> rgba;
    uint32_t width = 0, height = 0;
};

inline long wrap_index(long i, long n)
{
    const long m = i % n;
    return m < 0 ? m + n : m;
}

inline f4 texture_lookup(const Texture& t, float u, float v)
{
    const double N = double(t.width), M = double(t.height);
    // wrap addressing: only the fractional part of the normalised coordinate matters
    const double x = (double(u) - std::floor(double(u))) * N - 0.5;
    const double y = (double(v) - std::floor(double(v))) * M - 0.5;
    const double fx = std::floor(x), fy = std::floor(y);
    // 1.8 fixed-point filter weights
    const double a = std::nearbyint((x - fx) * 256.0) / 256.0;
    const double b = std::nearbyint((y - fy) * 256.0) / 256.0;
    const long i0 = wrap_index(long(fx), long(t.width)), i1 = wrap_index(long(fx) + 1, long(t.width));
    const long j0 = wrap_index(long(fy), long(t.height)), j1 = wrap_index(long(fy) + 1, long(t.height));
    auto texel = [&](long i, long j, int c) { return double(t.rgba[4 * (size_t(j) * t.width + size_t(i)) + c]) / 255.0; };
    float out[4];
    for (int c = 0; c < 4; ++c)
        out[c] = float((1.0 - a) * (1.0 - b) * texel(i0, j0, c) + a * (1.0 - b) * texel(i1, j0, c) + (1.0 - a) * b * texel(i0, j1, c) +
                       a * b * texel(i1, j1, c));
    return f4{ out[0], out[1], out[2], out[3] };
}

} // namespace orc
