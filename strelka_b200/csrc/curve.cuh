// Cubic B-spline curve segments on the device.
//
// (1) Evaluation + shading normal/tangent exactly as the reference's closest-hit program uses them
//     (cuda/curve.h: CubicInterpolator::initializeFromBSpline :177-187, position4 :231-234, velocity4
//     :246-254, acceleration4 :266-269, surfaceNormal<.,2> :306-353, curveTangent :412-417).
// (2) This backend's own ray / round-curve intersector.  The reference has none: it relies on OptiX's
//     closed built-in ROUND_CUBIC_BSPLINE primitive (OptixRender.cpp:553-560) -> parity unpinned; the
//     CPU oracle holds the same float solver (for image parity) plus a slow double-precision bracketing
//     solver that validates it (tests/test_curves.py).
//
// Surface definition: union of spheres |x - c(u)| <= r(u), u in [0,1], no end caps.  With a unit ray
// direction d, z(u) = (c(u)-o).d, rho(u) = distance of c(u) to the ray axis and g = rho^2 - r^2, the
// ray enters the sphere at parameter u at s(u) = z - sqrt(-g); the hit is the interior stationary
// point of s.  ds/du = 0  <=>  G(u) = 2 z' sqrt(-g) + g' = 0.  The solver iterates on a local model
// (g quadratic, z' constant) whose root is closed-form; where the model has no root (ray outside the
// tube at this u) it degrades to a Newton step towards the closest approach argmin g.
#pragma once
#include "hd.cuh"

#ifndef SB_CURVE_EARLY_REJECT
#define SB_CURVE_EARLY_REJECT 1
#endif

namespace sb
{

struct CubicSeg
{
    float4 p[4]; // P(u) = p0 u^3 + p1 u^2 + p2 u + p3 (xyz, w = radius)
};
SB_HD CubicSeg cubic_from_bspline(const float4 q[4])
{
    CubicSeg c;
    c.p[0] = (q[0] * (-1.0f) + q[1] * (3.0f) + q[2] * (-3.0f) + q[3]) / 6.0f;
    c.p[1] = (q[0] * (3.0f) + q[1] * (-6.0f) + q[2] * (3.0f)) / 6.0f;
    c.p[2] = (q[0] * (-3.0f) + q[2] * (3.0f)) / 6.0f;
    c.p[3] = (q[0] * (1.0f) + q[1] * (4.0f) + q[2] * (1.0f)) / 6.0f;
    return c;
}
SB_HD float4 cubic_position4(const CubicSeg& c, float u)
{
    return (((c.p[0] * u) + c.p[1]) * u + c.p[2]) * u + c.p[3];
}
SB_HD float4 cubic_velocity4(const CubicSeg& c, float u)
{
    if (u == 0)
        u = 0.000001f;
    if (u == 1)
        u = 0.999999f;
    return ((3.0f * c.p[0] * u) + 2.0f * c.p[1]) * u + c.p[2];
}
SB_HD float4 cubic_acceleration4(const CubicSeg& c, float u)
{
    return 6.0f * c.p[0] * u + 2.0f * c.p[1];
}
SB_HD float3 cubic_surface_normal(const CubicSeg& bc, float u, float3& ps)
{
    float3 normal;
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
        const float4 p4 = cubic_position4(bc, u);
        const float3 p = mk3(p4);
        const float r = p4.w;
        const float4 d4 = cubic_velocity4(bc, u);
        const float3 d = mk3(d4);
        const float dr = d4.w;
        float dd = dot(d, d);
        float3 o1 = ps - p;
        o1 = o1 - (dot(o1, d) / dd) * d;
        o1 *= r / length(o1);
        ps = p + o1;
        dd -= dot(mk3(cubic_acceleration4(bc, u)), o1);
        normal = dd * o1 - (dr * r) * d;
    }
    return normalize(normal);
}
SB_HD float3 cubic_tangent(const CubicSeg& bc, float u)
{
    return normalize(mk3(cubic_velocity4(bc, u)));
}

// A traversal record holds the POWER-BASIS coefficients of one sub-span of a segment in world space:
// P(s) = c0 s^3 + c1 s^2 + c2 s + c3, s in [0,1], xyz + radius in w.  Splitting every segment into K spans
// gives the BVH K short, nearly straight pieces with tight boxes (a 1.5 cm hair segment that runs diagonally
// has an AABB hundreds of hair widths thick; K = 8 cut the primitives tested per ray by >10x on the 1 M-segment
// scene) and lets the solver below start from a good chord guess.
struct CurveSpan
{
    float4 c[4];
};
// span k of K of the B-spline segment with control points q (w = radius).  Fixed expression order: the CPU
// oracle evaluates the same tree so that both sides hold identical records.
SB_HD CurveSpan curve_span(const float4 q[4], uint32_t k, uint32_t K)
{
    const float s6 = 1.0f / 6.0f;
    const float4 a = (q[3] - q[0] + (q[1] - q[2]) * 3.0f) * s6;
    const float4 b = (q[0] + q[2]) * 0.5f - q[1];
    const float4 c = (q[2] - q[0]) * 0.5f;
    const float4 e = (q[0] + q[2] + q[1] * 4.0f) * s6;
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
SB_HD void curve_span_bounds(const CurveSpan& r, float3& lo, float3& hi)
{
    const float third = 1.0f / 3.0f;
    const float4 B0 = r.c[3];
    const float4 B1 = r.c[3] + r.c[2] * third;
    const float4 B2 = r.c[3] + r.c[2] * (2.0f * third) + r.c[1] * third;
    const float4 B3 = r.c[3] + r.c[2] + r.c[1] + r.c[0];
    const float rmax = fmaxf(fmaxf(fabsf(B0.w), fabsf(B1.w)), fmaxf(fabsf(B2.w), fabsf(B3.w))) * 1.0001f;
    lo = mk3(fminf(fminf(B0.x, B1.x), fminf(B2.x, B3.x)) - rmax, fminf(fminf(B0.y, B1.y), fminf(B2.y, B3.y)) - rmax,
             fminf(fminf(B0.z, B1.z), fminf(B2.z, B3.z)) - rmax);
    hi = mk3(fmaxf(fmaxf(B0.x, B1.x), fmaxf(B2.x, B3.x)) + rmax, fmaxf(fmaxf(B0.y, B1.y), fmaxf(B2.y, B3.y)) + rmax,
             fmaxf(fmaxf(B0.z, B1.z), fmaxf(B2.z, B3.z)) + rmax);
    // the power -> Bezier conversion rounds: pad by a few ulps of the coordinates
    const float pad = 4.0e-7f * fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fmaxf(fabsf(lo.y), fabsf(hi.y)), fmaxf(fabsf(lo.z), fabsf(hi.z))));
    lo = lo - mk3(pad);
    hi = hi + mk3(pad);
}

// ray vs one span of a round cubic curve; cf = power-basis coefficients (w = radius).  s in (0,1) on a hit.
SB_HD bool intersect_round_cubic(const float4 cf[4], const float3& o, const float3& dIn, float tmin, float tmax, float& tOut, float& uOut)
{
    const float dl2 = dot_fma(dIn, dIn);
    if (!(dl2 > 0.0f))
        return false;
    const float invLen = 1.0f / sqrtf(dl2);
    const float3 d = dIn * invLen;
    const float4 a4 = cf[0], b4 = cf[1], c4 = cf[2];
    float4 e4 = cf[3];
    e4.x -= o.x;
    e4.y -= o.y;
    e4.z -= o.z;
    // ray-centric frame (b1, b2, d): the curve becomes X(u), Y(u) across the ray, Z(u) along it, R(u).
    // Working with the perpendicular components avoids the |P|^2 - z^2 cancellation that would swamp
    // hair-thin radii a few metres from the ray origin.
    float3 b1, b2;
    {
        const float sg = copysignf(1.0f, d.z);
        const float k = -1.0f / (sg + d.z);
        const float m = d.x * d.y * k;
        b1 = mk3(1.0f + sg * d.x * d.x * k, sg * m, -sg * d.x);
        b2 = mk3(m, sg + d.y * d.y * k, -d.y);
    }
    const float3 a3 = mk3(a4), b3 = mk3(b4), c3 = mk3(c4), e3 = mk3(e4);
    const float ax = dot_fma(a3, b1), bx = dot_fma(b3, b1), cx = dot_fma(c3, b1), ex = dot_fma(e3, b1);
    const float ay = dot_fma(a3, b2), by = dot_fma(b3, b2), cy = dot_fma(c3, b2), ey = dot_fma(e3, b2);
    const float az = dot_fma(a3, d), bz = dot_fma(b3, d), cz = dot_fma(c3, d), ez = dot_fma(e3, d);
    const float ar = a4.w, br = b4.w, cr = c4.w, er = e4.w;
    // initial guess: closest approach of the ray axis to the chord P(0)P(1), in the 2-D cross-section
    const float Bx = ax + bx + cx, By = ay + by + cy;
    const float bb = fmaf(Bx, Bx, By * By);
#if SB_CURVE_EARLY_REJECT
    // Early out (conservative; nine tests in ten on a hair scene are misses that would otherwise run the iteration
    // to convergence): in the cross-section the ray axis is the origin, and the curve lies in the convex hull of its
    // Bezier points.  If the origin is farther from the chord LINE than the hull's largest excursion from that line
    // plus the largest radius, no point of the tube covers it.  All terms are scaled by |B|; 1e-3 relative and a few
    // ulps absolute of slack cover the roundings (the solver below is the hit definition; this test only skips it).
    {
        const float c0 = fmaf(ex, By, -(ey * Bx)); // cross(P(0), B)
        const float k1x = cx * (1.0f / 3.0f), k1y = cy * (1.0f / 3.0f); // Bezier point 1 - point 0
        const float k2x = fmaf(2.0f, k1x, bx * (1.0f / 3.0f)), k2y = fmaf(2.0f, k1y, by * (1.0f / 3.0f)); // point 2 - point 0
        const float d1 = fmaf(k1x, By, -(k1y * Bx)), d2 = fmaf(k2x, By, -(k2y * Bx)); // excursions of the inner points (x |B|)
        const float exc = fmaxf(fmaxf(fabsf(d1), fabsf(d2)), 0.0f);
        const float rmax = fmaxf(fmaxf(fabsf(er), fabsf(ar + br + cr + er)), fmaxf(fabsf(fmaf(cr, 1.0f / 3.0f, er)), fabsf(fmaf(2.0f / 3.0f, cr, fmaf(br, 1.0f / 3.0f, er)))));
        const float reach = fmaf(rmax, sqrtf(bb), exc);
        if (fabsf(c0) > fmaf(reach, 1.001f, 1e-6f * (fabsf(ex * By) + fabsf(ey * Bx) + 1e-30f)))
            return false;
    }
#endif
    float u = (bb > 1e-30f) ? clampf(-fmaf(ex, Bx, ey * By) / bb, 0.0f, 1.0f) : 0.5f;
    for (int it = 0; it < 10; ++it)
    {
        const float X = fmaf(fmaf(fmaf(ax, u, bx), u, cx), u, ex), X1 = fmaf(fmaf(3.0f * ax, u, 2.0f * bx), u, cx), X2 = fmaf(6.0f * ax, u, 2.0f * bx);
        const float Y = fmaf(fmaf(fmaf(ay, u, by), u, cy), u, ey), Y1 = fmaf(fmaf(3.0f * ay, u, 2.0f * by), u, cy), Y2 = fmaf(6.0f * ay, u, 2.0f * by);
        const float R = fmaf(fmaf(fmaf(ar, u, br), u, cr), u, er), R1 = fmaf(fmaf(3.0f * ar, u, 2.0f * br), u, cr), R2 = fmaf(6.0f * ar, u, 2.0f * br);
        const float z1 = fmaf(fmaf(3.0f * az, u, 2.0f * bz), u, cz);
        const float g0 = fmaf(X, X, fmaf(Y, Y, -(R * R)));
        const float g1 = 2.0f * fmaf(X, X1, fmaf(Y, Y1, -(R * R1)));
        const float g2 = 2.0f * (fmaf(X1, X1, fmaf(X, X2, fmaf(Y1, Y1, Y * Y2))) - fmaf(R1, R1, R * R2));
        float delta;
        if (!(g2 > 0.0f))
        {
            delta = (g1 > 0.0f) ? -0.25f : 0.25f;
        }
        else
        {
            const float m = fmaf(g1, g1, -2.0f * g2 * g0); // model of g has real roots <=> the ray pierces the tube here
            if (m >= 0.0f)
            {
                const float zz = 2.0f * z1 * z1;
                const float disc = zz * m / (g2 + zz);
                delta = (-g1 - copysignf(sqrtf(disc), z1)) / g2;
            }
            else
            {
                delta = -g1 / g2;
            }
        }
        const float un = clampf(u + delta, 0.0f, 1.0f);
        const float step = fabsf(un - u);
        u = un;
        if (step < 2e-6f)
            break;
    }
    if (!(u > 0.0f && u < 1.0f))
        return false; // pinned at an end: that would be an end cap / belongs to the neighbouring segment
    const float X = fmaf(fmaf(fmaf(ax, u, bx), u, cx), u, ex);
    const float Y = fmaf(fmaf(fmaf(ay, u, by), u, cy), u, ey);
    const float Z = fmaf(fmaf(fmaf(az, u, bz), u, cz), u, ez);
    const float R = fmaf(fmaf(fmaf(ar, u, br), u, cr), u, er);
    const float g = fmaf(X, X, fmaf(Y, Y, -(R * R)));
    if (!(g < 0.0f))
        return false;
    const float sHit = Z - sqrtf(-g);
    const float t = sHit * invLen;
    if (!(t > tmin && t < tmax))
        return false;
    tOut = t;
    uOut = u;
    return true;
}

} // namespace sb
