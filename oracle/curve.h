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

// ------------------------------------------------------------------------------------------
// The single-precision solver that DEFINES the curve hit for image parity, and the span records it works on:
// the same definitions the CUDA kernels use (strelka_b200/csrc/curve.cuh), restated here.  Hair is a few tens
// of micrometres thick and metres away, so a float solver and the double bracketing solver above legitimately
// differ by a fraction of a percent of the radius -- enough to decorrelate individual paths.  The double solver
// is therefore kept as the VALIDATOR of this one (tests/test_curves.py: |dt| <= 0.05 r, same hit/miss away
// from silhouettes) while renders use this definition on both sides.
// ------------------------------------------------------------------------------------------
struct CurveSpan
{
    f4 c[4];
};
// span k of K of the B-spline segment with control points q (w = radius).  Fixed expression order: the CPU
// oracle evaluates the same tree so that both sides hold identical records.
inline CurveSpan curve_span(const f4 q[4], uint32_t k, uint32_t K)
{
    const float s6 = 1.0f / 6.0f;
    const f4 a = (q[3] - q[0] + (q[1] - q[2]) * 3.0f) * s6;
    const f4 b = (q[0] + q[2]) * 0.5f - q[1];
    const f4 c = (q[2] - q[0]) * 0.5f;
    const f4 e = (q[0] + q[2] + q[1] * 4.0f) * s6;
    const float h = 1.0f / float(K);
    const float u0 = float(k) * h;
    CurveSpan r;
    r.c[0] = a * (h * h * h);
    r.c[1] = (a * (3.0f * u0) + b) * (h * h);
    r.c[2] = ((a * (3.0f * u0) + b * 2.0f) * u0 + c) * h;
    r.c[3] = ((a * u0 + b) * u0 + c) * u0 + e;
    return r;
}
// conservative box: Bezier control points of the span (convex hull) grown by the largest Bezier radius weight
inline void curve_span_bounds(const CurveSpan& r, f3& lo, f3& hi)
{
    const float third = 1.0f / 3.0f;
    const f4 B0 = r.c[3];
    const f4 B1 = r.c[3] + r.c[2] * third;
    const f4 B2 = r.c[3] + r.c[2] * (2.0f * third) + r.c[1] * third;
    const f4 B3 = r.c[3] + r.c[2] + r.c[1] + r.c[0];
    const float rmax = std::fmax(std::fmax(std::fabs(B0.w), std::fabs(B1.w)), std::fmax(std::fabs(B2.w), std::fabs(B3.w))) * 1.0001f;
    lo = f3{ std::fmin(std::fmin(B0.x, B1.x), std::fmin(B2.x, B3.x)) - rmax, std::fmin(std::fmin(B0.y, B1.y), std::fmin(B2.y, B3.y)) - rmax,
             std::fmin(std::fmin(B0.z, B1.z), std::fmin(B2.z, B3.z)) - rmax };
    hi = mk3(std::fmax(std::fmax(B0.x, B1.x), std::fmax(B2.x, B3.x)) + rmax, std::fmax(std::fmax(B0.y, B1.y), std::fmax(B2.y, B3.y)) + rmax,
             std::fmax(std::fmax(B0.z, B1.z), std::fmax(B2.z, B3.z)) + rmax);
    // the power -> Bezier conversion rounds: pad by a few ulps of the coordinates
    const float pad = 4.0e-7f * std::fmax(std::fmax(std::fabs(lo.x), std::fabs(hi.x)), std::fmax(std::fmax(std::fabs(lo.y), std::fabs(hi.y)), std::fmax(std::fabs(lo.z), std::fabs(hi.z))));
    lo = lo - mk3(pad);
    hi = hi + mk3(pad);
}

// ray vs one span of a round cubic curve; cf = power-basis coefficients (w = radius).  s in (0,1) on a hit.
inline bool intersect_round_cubic_f32(const f4 cf[4], const f3& o, const f3& dIn, float tmin, float tmax, float& tOut, float& uOut)
{
    const float dl2 = dot_fma(dIn, dIn);
    if (!(dl2 > 0.0f))
        return false;
    const float invLen = 1.0f / std::sqrt(dl2);
    const f3 d = dIn * invLen;
    const f4 a4 = cf[0], b4 = cf[1], c4 = cf[2];
    f4 e4 = cf[3];
    e4.x -= o.x;
    e4.y -= o.y;
    e4.z -= o.z;
    // ray-centric frame (b1, b2, d): the curve becomes X(u), Y(u) across the ray, Z(u) along it, R(u).
    // Working with the perpendicular components avoids the |P|^2 - z^2 cancellation that would swamp
    // hair-thin radii a few metres from the ray origin.
    f3 b1, b2;
    {
        const float sg = std::copysign(1.0f, d.z);
        const float k = -1.0f / (sg + d.z);
        const float m = d.x * d.y * k;
        b1 = f3{ 1.0f + sg * d.x * d.x * k, sg * m, -sg * d.x };
        b2 = f3{ m, sg + d.y * d.y * k, -d.y };
    }
    const f3 a3 = mk3(a4), b3 = mk3(b4), c3 = mk3(c4), e3 = mk3(e4);
    const float ax = dot_fma(a3, b1), bx = dot_fma(b3, b1), cx = dot_fma(c3, b1), ex = dot_fma(e3, b1);
    const float ay = dot_fma(a3, b2), by = dot_fma(b3, b2), cy = dot_fma(c3, b2), ey = dot_fma(e3, b2);
    const float az = dot_fma(a3, d), bz = dot_fma(b3, d), cz = dot_fma(c3, d), ez = dot_fma(e3, d);
    const float ar = a4.w, br = b4.w, cr = c4.w, er = e4.w;
    // initial guess: closest approach of the ray axis to the chord P(0)P(1), in the 2-D cross-section
    const float Bx = ax + bx + cx, By = ay + by + cy;
    const float bb = std::fmaf(Bx, Bx, By * By);
    float u = (bb > 1e-30f) ? clampf(-std::fmaf(ex, Bx, ey * By) / bb, 0.0f, 1.0f) : 0.5f;
    for (int it = 0; it < 10; ++it)
    {
        const float X = std::fmaf(std::fmaf(std::fmaf(ax, u, bx), u, cx), u, ex), X1 = std::fmaf(std::fmaf(3.0f * ax, u, 2.0f * bx), u, cx), X2 = std::fmaf(6.0f * ax, u, 2.0f * bx);
        const float Y = std::fmaf(std::fmaf(std::fmaf(ay, u, by), u, cy), u, ey), Y1 = std::fmaf(std::fmaf(3.0f * ay, u, 2.0f * by), u, cy), Y2 = std::fmaf(6.0f * ay, u, 2.0f * by);
        const float R = std::fmaf(std::fmaf(std::fmaf(ar, u, br), u, cr), u, er), R1 = std::fmaf(std::fmaf(3.0f * ar, u, 2.0f * br), u, cr), R2 = std::fmaf(6.0f * ar, u, 2.0f * br);
        const float z1 = std::fmaf(std::fmaf(3.0f * az, u, 2.0f * bz), u, cz);
        const float g0 = std::fmaf(X, X, std::fmaf(Y, Y, -(R * R)));
        const float g1 = 2.0f * std::fmaf(X, X1, std::fmaf(Y, Y1, -(R * R1)));
        const float g2 = 2.0f * (std::fmaf(X1, X1, std::fmaf(X, X2, std::fmaf(Y1, Y1, Y * Y2))) - std::fmaf(R1, R1, R * R2));
        float delta;
        if (!(g2 > 0.0f))
        {
            delta = (g1 > 0.0f) ? -0.25f : 0.25f;
        }
        else
        {
            const float m = std::fmaf(g1, g1, -2.0f * g2 * g0); // model of g has real roots <=> the ray pierces the tube here
            if (m >= 0.0f)
            {
                const float zz = 2.0f * z1 * z1;
                const float disc = zz * m / (g2 + zz);
                delta = (-g1 - std::copysign(std::sqrt(disc), z1)) / g2;
            }
            else
            {
                delta = -g1 / g2;
            }
        }
        const float un = clampf(u + delta, 0.0f, 1.0f);
        const float step = std::fabs(un - u);
        u = un;
        if (step < 2e-6f)
            break;
    }
    if (!(u > 0.0f && u < 1.0f))
        return false; // pinned at an end: that would be an end cap / belongs to the neighbouring segment
    const float X = std::fmaf(std::fmaf(std::fmaf(ax, u, bx), u, cx), u, ex);
    const float Y = std::fmaf(std::fmaf(std::fmaf(ay, u, by), u, cy), u, ey);
    const float Z = std::fmaf(std::fmaf(std::fmaf(az, u, bz), u, cz), u, ez);
    const float R = std::fmaf(std::fmaf(std::fmaf(ar, u, br), u, cr), u, er);
    const float g = std::fmaf(X, X, std::fmaf(Y, Y, -(R * R)));
    if (!(g < 0.0f))
        return false;
    const float sHit = Z - std::sqrt(-g);
    const float t = sHit * invLen;
    if (!(t > tmin && t < tmax))
        return false;
    tOut = t;
    uOut = u;
    return true;
}


} // namespace orc
