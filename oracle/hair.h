// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Hair fibre scattering of SB_MATERIAL_HAIR, written from the publications and NOT from the device file
// (strelka_b200/csrc/hair.cuh): double precision, libstdc++'s std::cyl_bessel_i for the Bessel function, the closed
// forms of the papers evaluated as written there.
//   [C16] Chiang, Bitterli, Tappan, Burley: A Practical and Controllable Hair and Fur Model for Production Path Tracing,
//         Eurographics 2016 (roughness -> variance / logistic scale fits, eqs. 7-8; cuticle tilts alpha, 2 alpha, 4 alpha)
//   [dE11] d'Eon, Francois, Hill, Letteri, Aubry: An Energy-Conserving Hair Reflectance Model, EGSR 2011
//         (longitudinal function M_p, eq. 7; attenuations A_p, eqs. 10-14; lobe centres Phi(p, h))
//   [P16] Pharr: The Implementation of a Hair Scattering Model (pbrt-v3 supplement 2016): trimmed logistic azimuthal
//         function, the lobe-sampling scheme (choose p by the luminance of A_p, exact M_p sampling, logistic N_p)
// This is MDL's df::chiang_hair_bsdf (the reference's hair materials, mdlPtxCodeGen.cpp:143-155) with
// diffuse_reflection_weight = 0; the reference fixes the azimuthal offset h = 0 through its constant texture coordinate
// (closest_hit.cu:446).  The MDL SDK itself is closed: PARITY UNPINNED against the reference, pinned by the
// normalisation / energy / histogram tests of tests/test_bsdf_pins.py.
#pragma once
#include "vec.h"
#include <cmath>

namespace orc
{
namespace hair
{

constexpr double kPiD = 3.14159265358979323846;
constexpr int kPMax = 3; // R, TT, TRT, residual

inline double log_bessel_i0(double x)
{
    if (x < 500.0)
        return std::log(std::cyl_bessel_i(0.0, x));
    // Abramowitz & Stegun 9.7.1
    const double r = 1.0 / x;
    return x - 0.5 * std::log(2.0 * kPiD * x) + std::log(1.0 + r * (1.0 / 8.0 + r * (9.0 / 128.0 + r * (225.0 / 3072.0))));
}

// [dE11] eq. 7: M_p(theta_i, theta_o) = exp(-sin_i sin_o / v) I0(cos_i cos_o / v) / (2 v sinh(1/v))
inline double longitudinal(double sinI, double cosI, double sinO, double cosO, double v)
{
    const double logM = log_bessel_i0(cosI * cosO / v) - sinI * sinO / v - 1.0 / v - std::log(v) - std::log1p(-std::exp(-2.0 / v));
    return std::exp(logM);
}

// exact unpolarised Fresnel reflectance, air -> dielectric
inline double fresnel(double cosI, double eta)
{
    cosI = std::min(std::max(cosI, 0.0), 1.0);
    const double s2 = (1.0 - cosI * cosI) / (eta * eta);
    if (s2 >= 1.0)
        return 1.0;
    const double cosT = std::sqrt(1.0 - s2);
    const double rs = (cosI - eta * cosT) / (cosI + eta * cosT);
    const double rp = (eta * cosI - cosT) / (eta * cosI + cosT);
    return 0.5 * (rs * rs + rp * rp);
}

inline double logistic_pdf(double x, double s)
{
    const double e = std::exp(-std::fabs(x) / s);
    return e / (s * (1.0 + e) * (1.0 + e));
}
inline double logistic_cdf(double x, double s)
{
    return 1.0 / (1.0 + std::exp(-x / s));
}

struct Model
{
    double v[kPMax + 1];
    double s;
    double tilt[3]; // alpha, 2 alpha, 4 alpha
    double eta;
    double sigma[3];
    double h = 0.0;

    explicit Model(const sb_material& m)
    {
        const double bm = std::min(std::max(double(m.hair_roughness_lon), 0.01), 1.0);
        const double bn = std::min(std::max(double(m.hair_roughness_azi), 0.01), 1.0);
        // [C16] eq. 7 / 8
        const double sq = 0.726 * bm + 0.812 * bm * bm + 3.7 * std::pow(bm, 20.0);
        v[0] = sq * sq;
        v[1] = 0.25 * v[0];
        v[2] = 4.0 * v[0];
        v[3] = v[2];
        s = std::sqrt(kPiD / 8.0) * (0.265 * bn + 1.194 * bn * bn + 5.372 * std::pow(bn, 22.0));
        // float sine of the float angle, like any consumer of the material record would form it; then exact doubling
        const double a = std::asin(double(std::sin(m.hair_cuticle_angle)));
        tilt[0] = a;
        tilt[1] = 2.0 * a;
        tilt[2] = 4.0 * a;
        eta = m.ior > 1.0f ? double(m.ior) : 1.55;
        for (int c = 0; c < 3; ++c)
            sigma[c] = std::max(double(m.hair_absorption[c]), 0.0);
    }

    // tilted outgoing elevation per lobe ([C16] section 3.3: R by -2 alpha, TT by +alpha, TRT by +4 alpha)
    double tilted_theta(int p, double thetaO) const
    {
        if (p == 0)
            return thetaO - tilt[1];
        if (p == 1)
            return thetaO + tilt[0];
        if (p == 2)
            return thetaO + tilt[2];
        return thetaO;
    }

    struct Lobes
    {
        double A[kPMax + 1][3];
        double prob[kPMax + 1];
        double gammaT;
    };

    Lobes attenuations(double sinO, double cosO) const
    {
        Lobes L;
        // [dE11] section 3.1 / [P16]: modified index, refracted offset, one internal chord
        const double etap = std::sqrt(std::max(eta * eta - sinO * sinO, 0.0)) / std::max(cosO, 1e-6);
        const double sinGT = std::min(std::max(h / etap, -1.0), 1.0);
        const double cosGT = std::sqrt(1.0 - sinGT * sinGT);
        L.gammaT = std::asin(sinGT);
        const double sinT = sinO / eta;
        const double cosT = std::sqrt(std::max(1.0 - sinT * sinT, 0.0));
        const double len = 2.0 * cosGT / std::max(cosT, 1e-6);
        const double f = fresnel(cosO * std::sqrt(std::max(1.0 - h * h, 0.0)), eta);
        double sum = 0.0;
        for (int c = 0; c < 3; ++c)
        {
            const double T = std::exp(-sigma[c] * len);
            L.A[0][c] = f;
            L.A[1][c] = (1.0 - f) * (1.0 - f) * T;
            L.A[2][c] = L.A[1][c] * T * f;
            L.A[3][c] = L.A[2][c] * T * f / (1.0 - T * f); // geometric series of all further internal reflections
        }
        for (int p = 0; p <= kPMax; ++p)
        {
            L.prob[p] = 0.299 * L.A[p][0] + 0.587 * L.A[p][1] + 0.114 * L.A[p][2];
            sum += L.prob[p];
        }
        for (int p = 0; p <= kPMax; ++p)
            L.prob[p] = sum > 0.0 ? L.prob[p] / sum : 0.0;
        return L;
    }

    double azimuthal(double phi, int p, double gammaT) const
    {
        const double gammaO = std::asin(h);
        double d = phi - (2.0 * p * gammaT - 2.0 * gammaO + p * kPiD);
        d = std::remainder(d, 2.0 * kPiD); // to [-pi, pi]
        return logistic_pdf(d, s) / (logistic_cdf(kPiD, s) - logistic_cdf(-kPiD, s));
    }

    // f * |cos| (rgb) and the sampling density for local directions (x along the fibre)
    void eval(const double wo[3], const double wi[3], double fcos[3], double& pdf) const
    {
        const double sinO = std::min(std::max(wo[0], -1.0), 1.0), cosO = std::sqrt(1.0 - sinO * sinO);
        const double sinI = std::min(std::max(wi[0], -1.0), 1.0), cosI = std::sqrt(1.0 - sinI * sinI);
        const double phi = std::atan2(wi[2], wi[1]) - std::atan2(wo[2], wo[1]);
        const double thetaO = std::asin(sinO);
        const Lobes L = attenuations(sinO, cosO);
        fcos[0] = fcos[1] = fcos[2] = 0.0;
        pdf = 0.0;
        for (int p = 0; p <= kPMax; ++p)
        {
            double w;
            if (p < kPMax)
            {
                const double th = tilted_theta(p, thetaO);
                w = longitudinal(sinI, cosI, std::sin(th), std::fabs(std::cos(th)), v[p]) * azimuthal(phi, p, L.gammaT);
            }
            else
            {
                w = longitudinal(sinI, cosI, sinO, cosO, v[p]) / (2.0 * kPiD);
            }
            for (int c = 0; c < 3; ++c)
                fcos[c] += L.A[p][c] * w;
            pdf += L.prob[p] * w;
        }
    }

    // [P16] sampling: u[2] picks the lobe, (u[0], u[1]) the elevation, u[3] the azimuth
    void sample(const double wo[3], const double u[4], double wi[3]) const
    {
        const double sinO = std::min(std::max(wo[0], -1.0), 1.0), cosO = std::sqrt(1.0 - sinO * sinO);
        const double thetaO = std::asin(sinO);
        const Lobes L = attenuations(sinO, cosO);
        int p = 0;
        double r = u[2];
        while (p < kPMax && r >= L.prob[p])
        {
            r -= L.prob[p];
            ++p;
        }
        const double th = tilted_theta(p, thetaO);
        const double sinOp = std::sin(th), cosOp = std::fabs(std::cos(th));
        const double u1 = std::max(u[0], 1e-5);
        const double c = 1.0 + v[p] * std::log(u1 + (1.0 - u1) * std::exp(-2.0 / v[p]));
        const double sn = std::sqrt(std::max(1.0 - c * c, 0.0));
        double sinI = -c * sinOp + sn * std::cos(2.0 * kPiD * u[1]) * cosOp;
        sinI = std::min(std::max(sinI, -1.0), 1.0);
        const double cosI = std::sqrt(1.0 - sinI * sinI);
        double dphi;
        if (p < kPMax)
        {
            const double lo = logistic_cdf(-kPiD, s), k = logistic_cdf(kPiD, s) - lo;
            double x = -s * std::log(1.0 / (u[3] * k + lo) - 1.0);
            x = std::min(std::max(x, -kPiD), kPiD);
            dphi = 2.0 * p * L.gammaT - 2.0 * std::asin(h) + p * kPiD + x;
        }
        else
        {
            dphi = 2.0 * kPiD * u[3];
        }
        const double phiI = std::atan2(wo[2], wo[1]) + dphi;
        wi[0] = sinI;
        wi[1] = cosI * std::cos(phiI);
        wi[2] = cosI * std::sin(phiI);
    }
};

// fibre frame from the shading inputs of the closest-hit program (tangent_u = curveTangent, normal = surfaceNormal)
struct Frame
{
    double t[3], n[3], b[3];
    Frame(const f3& tangent, const f3& normal)
    {
        const double tl = std::sqrt(double(tangent.x) * tangent.x + double(tangent.y) * tangent.y + double(tangent.z) * tangent.z);
        t[0] = tangent.x / tl;
        t[1] = tangent.y / tl;
        t[2] = tangent.z / tl;
        const double d = normal.x * t[0] + normal.y * t[1] + normal.z * t[2];
        n[0] = normal.x - d * t[0];
        n[1] = normal.y - d * t[1];
        n[2] = normal.z - d * t[2];
        double nl = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (!(nl > 1e-6))
        {
            const double a[3] = { std::fabs(t[0]) < 0.9 ? 1.0 : 0.0, std::fabs(t[0]) < 0.9 ? 0.0 : 1.0, 0.0 };
            const double da = a[0] * t[0] + a[1] * t[1] + a[2] * t[2];
            for (int i = 0; i < 3; ++i)
                n[i] = a[i] - da * t[i];
            nl = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        }
        for (int i = 0; i < 3; ++i)
            n[i] /= nl;
        b[0] = t[1] * n[2] - t[2] * n[1];
        b[1] = t[2] * n[0] - t[0] * n[2];
        b[2] = t[0] * n[1] - t[1] * n[0];
    }
    void to_local(const f3& w, double out[3]) const
    {
        out[0] = w.x * t[0] + w.y * t[1] + w.z * t[2];
        out[1] = w.x * n[0] + w.y * n[1] + w.z * n[2];
        out[2] = w.x * b[0] + w.y * b[1] + w.z * b[2];
    }
    f3 to_world(const double w[3]) const
    {
        const double x = w[0] * t[0] + w[1] * n[0] + w[2] * b[0], y = w[0] * t[1] + w[1] * n[1] + w[2] * b[1],
                     z = w[0] * t[2] + w[1] * n[2] + w[2] * b[2];
        const double l = std::sqrt(x * x + y * y + z * z);
        return f3{ float(x / l), float(y / l), float(z / l) };
    }
};

} // namespace hair
} // namespace orc
