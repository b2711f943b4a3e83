"""Independent numpy statements of the material models, written from the publications -- NOT from
strelka_b200/csrc/bsdf.cuh, hair.cuh or oracle/bsdf.h, hair.h -- used to pin those implementations
(tests/test_bsdf_pins.py on the CPU oracle, tests/test_gpu_bsdf_pins.py on the device).

  Lambert                      f = rho / pi
  GGX microfacet BRDF          Walter, Marschner, Li, Torrance: Microfacet Models for Refraction through Rough Surfaces
                               (EGSR 2007): f = F D G / (4 |n.i| |n.o|), D eq. 33, G1 eq. 34, G = G1(i) G1(o)
  visible-normal density       Heitz: Sampling the GGX Distribution of Visible Normals (JCGT 2018), eq. 1, 3:
                               D_v(m) = G1(v) max(0, v.m) D(m) / (n.v);  pdf(o) = D_v / (4 v.m)
  Schlick Fresnel              F = F0 + (1 - F0)(1 - cos)^5
  UsdPreviewSurface layering   Pixar UsdPreviewSurface specification: metallic workflow F0 = mix(((1-ior)/(1+ior))^2, base,
                               metallic), diffuse = base (1 - metallic); specular workflow F0 = specularColor; clearcoat = GGX
                               with F0 0.04 on top; layering weights as DESIGN.md section 6 states them
  hair fibre                   d'Eon et al. 2011 (M_p, A_p), Chiang et al. 2016 (roughness fits, tilts), Pharr 2016 (trimmed
                               logistic N_p); h = 0

Everything is float64 and vectorised over the last axis of directions.  Directions are in a local frame with the shading
normal = +z (surfaces) or the fibre tangent = +x, normal = +y (hair).
"""
import numpy as np
from scipy.special import ive

PI = np.pi


def schlick(f0, c):
    return f0 + (1.0 - f0) * (1.0 - c) ** 5


def ggx_D(alpha, cos_m):
    """Walter 2007 eq. 33 with tan^2 = (1 - cos^2) / cos^2"""
    c2 = cos_m * cos_m
    t2 = (1.0 - c2) / np.maximum(c2, 1e-300)
    return np.where(cos_m > 0, alpha**2 / (PI * c2 * c2 * (alpha**2 + t2) ** 2), 0.0)


def ggx_G1(alpha, cos_v):
    """Walter 2007 eq. 34"""
    c2 = cos_v * cos_v
    t2 = (1.0 - c2) / np.maximum(c2, 1e-300)
    return np.where(cos_v > 0, 2.0 / (1.0 + np.sqrt(1.0 + alpha**2 * t2)), 0.0)


def luminance(c):
    c = np.asarray(c, dtype=np.float64)
    return 0.299 * c[..., 0] + 0.587 * c[..., 1] + 0.114 * c[..., 2]


class PreviewSurface:
    """UsdPreviewSurface as this backend defines it, from the formulas above."""

    def __init__(self, base_color, roughness, metallic, ior=1.5, clearcoat=0.0, clearcoat_roughness=0.01, specular_color=(0, 0, 0),
                 use_specular_workflow=False):
        base = np.asarray(base_color, dtype=np.float64)
        metallic = float(np.clip(metallic, 0, 1))
        self.alpha = max(float(np.clip(roughness, 0, 1)) ** 2, 1e-3)
        f0d = ((1.0 - ior) / (1.0 + ior)) ** 2
        if use_specular_workflow:
            self.F0 = np.asarray(specular_color, dtype=np.float64)
            self.diff = base
        else:
            self.F0 = f0d * (1.0 - metallic) + base * metallic
            self.diff = base * (1.0 - metallic)
        self.cc = float(np.clip(clearcoat, 0, 1))
        self.cc_alpha = max(float(np.clip(clearcoat_roughness, 0, 1)) ** 2, 1e-3)

    def layer_weights(self, cos1):
        """attenuation by the coat, diffuse weight under the specular layer, lobe selection probabilities"""
        fc = self.cc * schlick(0.04, cos1)
        att = 1.0 - fc
        wd = 1.0 - schlick(self.F0.mean(), cos1)
        ws = att * luminance(schlick(self.F0, cos1))
        wdl = att * wd * luminance(self.diff)
        tot = fc + ws + wdl
        return att, wd, fc / tot, ws / tot, wdl / tot

    def eval(self, k1, k2):
        """f * cos(k2) split into (diffuse, glossy) and the pdf of the one-sample lobe mixture; k1 (3,), k2 (..., 3)"""
        k1 = np.asarray(k1, dtype=np.float64)
        k2 = np.asarray(k2, dtype=np.float64)
        c1, c2 = k1[2], k2[..., 2]
        up = (c2 > 0) & (c1 > 0)
        att, wd, pc, ps, pd = self.layer_weights(c1)
        h = k1 + k2
        h = h / np.linalg.norm(h, axis=-1, keepdims=True)
        ch = np.clip(h[..., 2], 0, 1)  # n.h
        vh = np.clip((h * k1).sum(-1), 0, 1)  # k1.h
        c2s = np.where(up, c2, 1.0)
        D, G = ggx_D(self.alpha, ch), ggx_G1(self.alpha, c1) * ggx_G1(self.alpha, c2s)
        spec = schlick(self.F0, vh[..., None]) * (att * D * G / (4.0 * c1 * c2s))[..., None]
        pdf = ps * ggx_G1(self.alpha, c1) * D / (4.0 * c1) + pd * c2s / PI
        if self.cc > 0:
            Dc, Gc = ggx_D(self.cc_alpha, ch), ggx_G1(self.cc_alpha, c1) * ggx_G1(self.cc_alpha, c2s)
            spec = spec + (self.cc * schlick(0.04, vh) * Dc * Gc / (4.0 * c1 * c2s))[..., None]
            pdf = pdf + pc * ggx_G1(self.cc_alpha, c1) * Dc / (4.0 * c1)
        glossy = spec * c2s[..., None]
        diffuse = self.diff * (att * wd * c2s / PI)[..., None]
        z = np.zeros_like(glossy)
        return np.where(up[..., None], diffuse, z), np.where(up[..., None], glossy, z), np.where(up, pdf, 0.0)


class Lambert:
    def __init__(self, color):
        self.color = np.asarray(color, dtype=np.float64)

    def eval(self, k1, k2):
        k2 = np.asarray(k2, dtype=np.float64)
        up = (k2[..., 2] > 0) & (k1[2] > 0)
        c = np.where(up, k2[..., 2], 0.0)
        d = self.color * (c / PI)[..., None]
        return d, np.zeros_like(d), c / PI


class Hair:
    """Chiang 2016 / d'Eon 2011 fibre scattering with azimuthal offset h = 0; local frame: x = tangent, y = normal."""

    def __init__(self, sigma_a, beta_m, beta_n, alpha, eta):
        bm, bn = float(np.clip(beta_m, 0.01, 1)), float(np.clip(beta_n, 0.01, 1))
        v0 = (0.726 * bm + 0.812 * bm**2 + 3.7 * bm**20) ** 2
        self.v = [v0, 0.25 * v0, 4.0 * v0, 4.0 * v0]
        self.s = np.sqrt(PI / 8.0) * (0.265 * bn + 1.194 * bn**2 + 5.372 * bn**22)
        self.alpha = float(np.arcsin(np.float64(np.sin(np.float32(alpha)))))
        self.eta = float(eta)
        self.sigma = np.asarray(sigma_a, dtype=np.float64)

    @staticmethod
    def M(sin_i, cos_i, sin_o, cos_o, v):
        # exp(-b) I0(a) / (2 v sinh(1/v)) with the exponentially scaled Bessel function ive(0, a) = I0(a) e^-a
        a, b = cos_i * cos_o / v, sin_i * sin_o / v
        return ive(0, a) * np.exp(a - b - 1.0 / v) / (v * (1.0 - np.exp(-2.0 / v)))

    @staticmethod
    def fresnel(c, eta):
        c = np.clip(c, 0, 1)
        s2 = (1 - c * c) / eta**2
        ct = np.sqrt(np.maximum(1 - s2, 0))
        rs = (c - eta * ct) / (c + eta * ct)
        rp = (eta * c - ct) / (eta * c + ct)
        return np.where(s2 >= 1, 1.0, 0.5 * (rs**2 + rp**2))

    def A(self, sin_o, cos_o):
        f = self.fresnel(cos_o, self.eta)  # h = 0: cos(gamma_o) = 1
        sin_t = sin_o / self.eta
        cos_t = np.sqrt(1 - sin_t**2)
        T = np.exp(-self.sigma * 2.0 / cos_t)  # h = 0: the chord through the axis, cos(gamma_t) = 1
        A = [np.full(3, f), (1 - f) ** 2 * T]
        A.append(A[1] * T * f)
        A.append(A[2] * T * f / (1 - T * f))
        return A

    def N(self, phi, p):
        d = np.remainder(phi - p * PI + PI, 2 * PI) - PI  # h = 0: Phi(p) = p pi
        e = np.exp(-np.abs(d) / self.s)
        cdf = lambda x: 1.0 / (1.0 + np.exp(-x / self.s))
        return e / (self.s * (1 + e) ** 2) / (cdf(PI) - cdf(-PI))

    def eval(self, wo, wi):
        """f * |cos| (..., 3) and the pdf of the lobe-mixture sampler; wo (3,), wi (..., 3)"""
        wo = np.asarray(wo, dtype=np.float64)
        wi = np.asarray(wi, dtype=np.float64)
        sin_o = np.clip(wo[0], -1, 1)
        cos_o = np.sqrt(1 - sin_o**2)
        th_o = np.arcsin(sin_o)
        sin_i = np.clip(wi[..., 0], -1, 1)
        cos_i = np.sqrt(1 - sin_i**2)
        phi = np.arctan2(wi[..., 2], wi[..., 1]) - np.arctan2(wo[2], wo[1])
        A = self.A(sin_o, cos_o)
        prob = np.array([luminance(a) for a in A])
        prob = prob / prob.sum()
        tilt = [th_o - 2 * self.alpha, th_o + self.alpha, th_o + 4 * self.alpha]
        f = np.zeros(wi.shape[:-1] + (3,))
        pdf = np.zeros(wi.shape[:-1])
        for p in range(3):
            w = self.M(sin_i, cos_i, np.sin(tilt[p]), abs(np.cos(tilt[p])), self.v[p]) * self.N(phi, p)
            f += A[p] * w[..., None]
            pdf += prob[p] * w
        w = self.M(sin_i, cos_i, sin_o, cos_o, self.v[3]) / (2 * PI)
        f += A[3] * w[..., None]
        pdf += prob[3] * w
        return f, pdf


# ---- quadrature helpers ---------------------------------------------------------------------------------------
def sphere_grid(n_theta, n_phi, pole=(0.0, 0.0, 1.0), hemisphere=False):
    """Midpoint rule in (cos theta, phi) about +z (exact solid-angle weights 2 * (2) pi / N).  Returns (dirs (N, 3), weight)."""
    lo = 0.0 if hemisphere else -1.0
    mu = lo + (np.arange(n_theta) + 0.5) * (1.0 - lo) / n_theta
    ph = (np.arange(n_phi) + 0.5) * 2 * PI / n_phi
    mu, ph = np.meshgrid(mu, ph, indexing="ij")
    st = np.sqrt(1 - mu * mu)
    d = np.stack([st * np.cos(ph), st * np.sin(ph), mu], axis=-1).reshape(-1, 3)
    return d, (1.0 - lo) * 2 * PI / (n_theta * n_phi)
