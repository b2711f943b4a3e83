// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Cubic B-spline segment evaluation and the curve shading normal, restating the pieces of
// cuda/curve.h the reference's closest-hit program uses (CubicInterpolator::initializeFromBSpline
// :177-187, position4 :231-234, velocity4 :246-254, acceleration4 :266-269, surfaceNormal<type 2>
// :306-353, curveTangent :412-417), plus this repo's own ray / round-cubic-curve intersector.
// The reference has no intersector of its own (it relies on OptiX's closed built-in
// ROUND_CUBIC_BSPLINE primitive, OptixRender.cpp:553-560) -> the intersector is "parity
// unpinned"; the oracle version below is a slow, robust double-precision solver.
#pragma once
#include "vec.h"
#include <vector>
#include <utility>

namespace orc
{

// P(u) = p0 u^3 + p1 u^2 + p2 u + p3, each with radius in .w
struct CubicSeg
{
    f4 p[4];
};

// B-spline control points q[0..3] -> polynomial coefficients (curve.h:177-187)
inline CubicSeg cubic_from_bspline(const f4 q[4])
{
    CubicSeg c;
    c.p[0] = (q[0] * (-1.0f) + q[1] * (3.0f) + q[2] * (-3.0f) + q[3]) / 6.0f;
    c.p[1] = (q[0] * (3.0f) + q[1] * (-6.0f) + q[2] * (3.0f)) / 6.0f;
    c.p[2] = (q[0] * (-3.0f) + q[2] * (3.0f)) / 6.0f;
    c.p[3] = (q[0] * (1.0f) + q[1] * (4.0f) + q[2] * (1.0f)) / 6.0f;
    return c;
}

inline f4 cubic_position4(const CubicSeg& c, float u)
{
    return (((c.p[0] * u) + c.p[1]) * u + c.p[2]) * u + c.p[3];
}
inline f4 cubic_velocity4(const CubicSeg& c, float u)
{
    // curve.h:246-254: avoid triple knots at the ends
    if (u == 0)
        u = 0.000001f;
    if (u == 1)
        u = 0.999999f;
    return ((3.0f * c.p[0] * u) + 2.0f * c.p[1]) * u + c.p[2];
}
inline f4 cubic_acceleration4(const CubicSeg& c, float u)
{
    return 6.0f * c.p[0] * u + 2.0f * c.p[1];
}

// surfaceNormal<CubicInterpolator, 2>, curve.h:306-353; ps is moved onto the surface.
inline f3 cubic_surface_normal(const CubicSeg& bc, float u, f3& ps)
{
    f3 normal;
    if (u == 0.0f)
    {
        normal = -mk3(cubic_velocity4(bc, 0));
    }
    else if (u == 1.0f)
    {
        normal = mk3(cubic_velocity4(bc, 1));
    }
    else
    {
        const f4 p4 = cubic_position4(bc, u);
        const f3 p = mk3(p4);
        const float r = p4.w;
        const f4 d4 = cubic_velocity4(bc, u);
        const f3 d = mk3(d4);
        const float dr = d4.w;
        float dd = dot(d, d);
        f3 o1 = ps - p;
        o1 = o1 - (dot(o1, d) / dd) * d;
        o1 *= r / length(o1);
        ps = p + o1;
        dd -= dot(mk3(cubic_acceleration4(bc, u)), o1);
        normal = dd * o1 - (dr * r) * d;
    }
    return normalize(normal);
}

// curveTangent, curve.h:412-417
inline f3 cubic_tangent(const CubicSeg& bc, float u)
{
    return normalize(mk3(cubic_velocity4(bc, u)));
}

// ------------------------------------------------------------------------------------------
// Ray vs round cubic curve segment (swept sphere of radius r(u) along c(u), u in [0,1], no end
// caps -- OptiX's default for cubic curves, OptixRender.cpp:282 leaves endcapFlags default).
//
// Definition used by BOTH the oracle and the CUDA kernel: the ray (unit direction d) enters the
// union of spheres {|x - c(u)| <= r(u)} at s* = min_u s(u), s(u) = z(u) - sqrt(r(u)^2 - rho(u)^2)
// where z(u) = (c(u)-o).d and rho(u) the distance from c(u) to the ray axis.  A hit is reported
// for an INTERIOR stationary point of s(u) (ds/du = 0, the envelope condition); minima pinned
// at u=0 / u=1 would lie on an end cap and are rejected.  Only entry points with s* in
// (tmin, tmax) count; a ray starting inside the curve reports nothing for that span.
//
// Oracle solver: double precision; g(u) = rho^2 - r^2 sampled on a fine grid to bracket each
// span where g < 0, bisection for the span ends, golden-section search for the minimum of s.
// ------------------------------------------------------------------------------------------
struct CurveHit
{
    bool hit;
    float t;
    float u;
};

inline CurveHit intersect_round_cubic(const f4 q[4], const f3& o, const f3& dir, float tmin, float tmax)
{
    // polynomial coefficients in double
    double P[4][4];
    for (int k = 0; k < 4; ++k)
    {
        const double q0 = (&q[0].x)[k], q1 = (&q[1].x)[k], q2 = (&q[2].x)[k], q3 = (&q[3].x)[k];
        P[0][k] = (-q0 + 3.0 * q1 - 3.0 * q2 + q3) / 6.0;
        P[1][k] = (3.0 * q0 - 6.0 * q1 + 3.0 * q2) / 6.0;
        P[2][k] = (-3.0 * q0 + 3.0 * q2) / 6.0;
        P[3][k] = (q0 + 4.0 * q1 + q2) / 6.0;
    }
    const double ox = o.x, oy = o.y, oz = o.z;
    double dx = dir.x, dy = dir.y, dz = dir.z;
    const double dl = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (!(dl > 0.0))
        return CurveHit{ false, 0, 0 };
    dx /= dl;
    dy /= dl;
    dz /= dl;

    struct Ev
    {
        double g, dg, z;
    };
    // g(u) = rho^2 - r^2 and its derivative
    auto eval = [&](double u) {
        double c[4], v[4];
        for (int k = 0; k < 4; ++k)
        {
            c[k] = ((P[0][k] * u + P[1][k]) * u + P[2][k]) * u + P[3][k];
            v[k] = (3.0 * P[0][k] * u + 2.0 * P[1][k]) * u + P[2][k];
        }
        const double wx = c[0] - ox, wy = c[1] - oy, wz = c[2] - oz;
        const double z = wx * dx + wy * dy + wz * dz;
        const double zp = v[0] * dx + v[1] * dy + v[2] * dz;
        const double rho2 = wx * wx + wy * wy + wz * wz - z * z;
        const double drho2 = 2.0 * (wx * v[0] + wy * v[1] + wz * v[2]) - 2.0 * z * zp;
        return Ev{ rho2 - c[3] * c[3], drho2 - 2.0 * c[3] * v[3], z };
    };
    auto sfun = [&](double u) {
        const Ev e = eval(u);
        return e.z - std::sqrt(std::max(0.0, -e.g));
    };

    // candidate abscissae: a uniform grid plus every local minimum of g between grid points
    const int N = 128;
    std::vector<std::pair<double, double>> cand; // (u, g)
    cand.reserve(2 * N + 2);
    Ev prev = eval(0.0);
    cand.emplace_back(0.0, prev.g);
    for (int k = 1; k <= N; ++k)
    {
        const double u1 = double(k) / N;
        const Ev cur = eval(u1);
        if (prev.dg < 0.0 && cur.dg >= 0.0)
        {
            double lo = double(k - 1) / N, hi = u1;
            for (int it = 0; it < 60; ++it)
            {
                const double mid = 0.5 * (lo + hi);
                if (eval(mid).dg < 0.0)
                    lo = mid;
                else
                    hi = mid;
            }
            const double um = 0.5 * (lo + hi);
            cand.emplace_back(um, eval(um).g);
        }
        cand.emplace_back(u1, cur.g);
        prev = cur;
    }
    auto root_between = [&](double uPos, double uNeg) {
        // g(uPos) >= 0, g(uNeg) < 0 -> point on the negative side of the root
        for (int it = 0; it < 70; ++it)
        {
            const double mid = 0.5 * (uPos + uNeg);
            if (eval(mid).g < 0.0)
                uNeg = mid;
            else
                uPos = mid;
        }
        return uNeg;
    };

    double best_s = 1e300, best_u = 0.0;
    bool found = false;
    const size_t M = cand.size();
    size_t k = 0;
    while (k < M)
    {
        if (!(cand[k].second < 0.0))
        {
            ++k;
            continue;
        }
        const size_t k0 = k;
        while (k + 1 < M && cand[k + 1].second < 0.0)
            ++k;
        const size_t k1 = k;
        const bool clip0 = (k0 == 0), clip1 = (k1 == M - 1);
        const double ua = clip0 ? 0.0 : root_between(cand[k0 - 1].first, cand[k0].first);
        const double ub = clip1 ? 1.0 : root_between(cand[k1 + 1].first, cand[k1].first);
        // minimise s on [ua, ub] by golden section (s is unimodal on a thin span)
        const double gr = 0.6180339887498949;
        double a = ua, b = ub;
        double c1 = b - gr * (b - a), c2 = a + gr * (b - a);
        double f1 = sfun(c1), f2 = sfun(c2);
        for (int it = 0; it < 100; ++it)
        {
            if (f1 < f2)
            {
                b = c2;
                c2 = c1;
                f2 = f1;
                c1 = b - gr * (b - a);
                f1 = sfun(c1);
            }
            else
            {
                a = c1;
                c1 = c2;
                f1 = f2;
                c2 = a + gr * (b - a);
                f2 = sfun(c2);
            }
        }
        const double um = 0.5 * (a + b);
        const double sm = sfun(um);
        const double tolu = 1e-6;
        const bool atStart = clip0 && (um - ua) < tolu;
        const bool atEnd = clip1 && (ub - um) < tolu;
        if (!atStart && !atEnd && sm > double(tmin) * dl && sm < double(tmax) * dl && sm < best_s)
        {
            best_s = sm;
            best_u = um;
            found = true;
        }
        k = k1 + 1;
    }
    if (!found)
        return CurveHit{ false, 0, 0 };
    // s was measured along the unit direction; convert to the caller's parametrisation
    return CurveHit{ true, float(best_s / dl), float(best_u) };
}

} // namespace orc
