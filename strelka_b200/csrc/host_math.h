// Small host-side computations of the render() prologue, shared by the C ABI and the CPU emulation
// harness: exposure (OptixRender.cpp:956-987), glm::inverse of the view matrix (OptixRender.cpp:953)
// and the clip-to-view matrix of Camera::updateAspectRatio -> perspective() (camera.cpp:61-131).
#pragma once
#include <algorithm>
#include <cmath>
#include "../../include/sb/sb_api.h"

namespace sb
{

// exposure, OptixRender.cpp:956-987 (white point (1,1,1))
inline void compute_exposure(const sb_settings& st, float out[3])
{
    float e[3] = { 1.0f / 1.0f, 1.0f / 1.0f, 1.0f / 1.0f };
    const float lum = e[0] * 0.299f + e[1] * 0.587f + e[2] * 0.114f;
    float k = st.cm2_factor;
    if (st.film_iso > 0.0f)
        k = st.cm2_factor * st.film_iso / (st.shutter_speed * st.f_stop * st.f_stop) / 100.0f;
    const float inv = 1.0f / lum;
    for (int i = 0; i < 3; ++i)
        out[i] = e[i] * k * inv;
}

// inverse of a general 4x4 (glm::inverse of the view matrix, OptixRender.cpp:953) in double
inline bool invert4(const double m[16], double inv[16])
{
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int col = 0; col < 4; ++col)
        {
            a[r][col] = m[r * 4 + col];
            a[r][4 + col] = (r == col) ? 1.0 : 0.0;
        }
    for (int col = 0; col < 4; ++col)
    {
        int piv = col;
        for (int r = col + 1; r < 4; ++r)
            if (std::fabs(a[r][col]) > std::fabs(a[piv][col]))
                piv = r;
        if (a[piv][col] == 0.0)
            return false;
        for (int k = 0; k < 8; ++k)
            std::swap(a[col][k], a[piv][k]);
        const double d = 1.0 / a[col][col];
        for (int k = 0; k < 8; ++k)
            a[col][k] *= d;
        for (int r = 0; r < 4; ++r)
        {
            if (r == col)
                continue;
            const double f = a[r][col];
            for (int k = 0; k < 8; ++k)
                a[r][k] -= f * a[col][k];
        }
    }
    for (int r = 0; r < 4; ++r)
        for (int col = 0; col < 4; ++col)
            inv[r * 4 + col] = a[r][4 + col];
    return true;
}

// Camera::updateAspectRatio -> perspective() (camera.cpp:61-131), uploaded transposed (OptixRender.cpp:954):
// rows (1/x,0,0,0), (0,1/y,0,0), (0,0,0,-1), (0,0,1/B,A/B).  The last row only feeds the unused w of the
// view-space point, so near/far do not influence rays; it is stored as (0,0,0,1).
inline void clip_to_view_from_fov(float fovYDeg, float aspect, float out[16])
{
    const float focal = 1.0f / std::tan((fovYDeg * 0.01745329251994329576923690768489f) / 2.0f);
    const float x = focal / aspect, y = focal;
    for (int i = 0; i < 16; ++i)
        out[i] = 0.0f;
    out[0] = 1 / x;
    out[5] = 1 / y;
    out[11] = -1.0f;
    out[15] = 1.0f;
}

// view (glm column-major) -> row-major viewToWorld = inverse(view)
inline bool view_to_world_from_view(const float view[16], float out[16])
{
    double m[16], inv[16];
    for (int r = 0; r < 4; ++r)
        for (int col = 0; col < 4; ++col)
            m[r * 4 + col] = view[col * 4 + r];
    if (!invert4(m, inv))
        return false;
    for (int i = 0; i < 16; ++i)
        out[i] = float(inv[i]);
    return true;
}

} // namespace sb
