"""Closed-form (host quadrature) expectation of the reference's direct-lighting estimator for ONE rect light over a
Lambert surface point -- independent of the oracle and of the device code.  It follows the reference's source lines:

  NEE (closest_hit.cu:260-324, 547-589):  sample a point on the light with pdf p_l (solid angle; uniform area sampling
      p_l = d^2 / (cos_l A), Lights.h:277-289, or spherical-rectangle sampling p_l = 1 / S, Lights.h:245-275), accept iff
      N.L > 0 and -L.n_l > 0; contribution  throughput * (Li * sat(N.L) / p_l) * mis(p_l, p_b) * bsdf,  bsdf = rho/pi * (N.L)
      (cosine INCLUDED: the double cosine of quirk Q4), p_b = (N.L) / pi, mis(a, b) = 1 / (1 + b/a) (Lights.h:28-31)
  emitter hit by the BSDF-sampled ray (OptixRender.cu:315-341): radiance += throughput * color * cos_l * mis(p_b, p_l')
      with p_l' = getLightPdf = d^2 / (cos_l A) / numLights WHATEVER the sampling method (Lights.h:200-243), throughput = rho
      (bsdf_over_pdf of the cosine-sampled Lambert lobe), cos_l = -dir.n_l (quirk Q5)

  E[depth-1 image] = Int_light  C rho/pi (N.L)^2 w_l                                  dOmega
  E[depth-2 image] = that  +  Int_light  C rho/pi (N.L) cos_l w_b                      dOmega,   dOmega = cos_l / d^2 dA
"""
import numpy as np


def spherical_rect_solid_angle(p0, e1, e2, x):
    """solid angle of the parallelogram p0 + s e1 + t e2 (a rectangle here) seen from x: sum of the two triangles'
    (Van Oosterom & Strackee 1983) -- not Urena's formula the implementations use"""
    def tri(a, b, c):
        a, b, c = a - x, b - x, c - x
        la, lb, lc = np.linalg.norm(a), np.linalg.norm(b), np.linalg.norm(c)
        num = np.dot(a, np.cross(b, c))
        den = la * lb * lc + np.dot(a, b) * lc + np.dot(a, c) * lb + np.dot(b, c) * la
        return abs(2.0 * np.arctan2(num, den))
    q0, q1, q2, q3 = p0, p0 + e1, p0 + e1 + e2, p0 + e2
    return tri(q0, q1, q2) + tri(q0, q2, q3)


def expected_radiance(light, x, n, rho, method, depth, n_quad=384):
    """light: numpy record (LIGHT_DTYPE); x: surface point; n: its normal; rho: albedo (3,); returns radiance (3,)"""
    P = np.asarray(light["points"], dtype=np.float64)[:, :3]
    C = np.asarray(light["color"], dtype=np.float64)[:3]
    p0, e1, e2 = P[0], P[1] - P[0], P[3] - P[0]
    cr = np.cross(e1, e2)
    A = np.linalg.norm(cr)
    nl = -cr / A
    s = (np.arange(n_quad) + 0.5) / n_quad
    su, sv = np.meshgrid(s, s, indexing="ij")
    y = p0 + su[..., None] * e1 + sv[..., None] * e2
    to = y - np.asarray(x, dtype=np.float64)
    d = np.linalg.norm(to, axis=-1)
    L = to / d[..., None]
    cos_n = L @ np.asarray(n, dtype=np.float64)
    cos_l = -(L @ nl)
    ok = (cos_n > 0) & (cos_l > 0)
    cos_n = np.where(ok, cos_n, 0.0)
    cos_ls = np.where(ok, cos_l, 1.0)
    d_omega = np.where(ok, cos_ls / d**2, 0.0) * (A / n_quad**2)
    p_area = d**2 / (cos_ls * A)  # getRectLightPdf, one light
    if method == 0:
        p_sample = p_area
    else:
        S = spherical_rect_solid_angle(p0, e1, e2, np.asarray(x, dtype=np.float64))
        p_sample = np.full_like(p_area, 1.0 / S)
    p_b = cos_n / np.pi
    w_l = 1.0 / (1.0 + p_b / p_sample)
    nee = (cos_n**2 * w_l * d_omega).sum()
    total = nee
    if depth >= 2:
        w_b = np.where(p_b > 0, 1.0 / (1.0 + p_area / np.maximum(p_b, 1e-300)), 0.0)
        total = total + (cos_n * cos_ls * w_b * d_omega).sum()
    return C * np.asarray(rho, dtype=np.float64) / np.pi * total
