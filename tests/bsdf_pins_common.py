"""Independent pins of the material protocol (closest_hit.cu:474-545): the same checks run against the CPU oracle
(tests/test_bsdf_pins.py) and against the device code through sb_test_bsdf (tests/test_gpu_bsdf_pins.py).

`run(material_record, packed_inputs (N, 19)) -> (sample (N, 8), eval (N, 7))` is the implementation under test;
the expectations come from tests/bsdf_ref.py (numpy, written from the publications) and from host quadrature.
"""
import numpy as np

import bsdf_ref as R
from strelka_b200 import _abi

PI = np.pi
EV_ABSORB, EV_DIFFUSE, EV_GLOSSY, EV_REFLECTION, EV_TRANSMISSION = 0, 1, 2, 8, 16


def material(model, **kw):
    m = np.zeros(1, dtype=_abi.MATERIAL_DTYPE)
    m["model"] = model
    m["base_color"] = kw.get("base_color", (1, 1, 1))
    m["roughness"] = kw.get("roughness", 0.5)
    m["metallic"] = kw.get("metallic", 0.0)
    m["ior"] = kw.get("ior", 1.5)
    m["opacity"] = 1.0
    m["clearcoat"] = kw.get("clearcoat", 0.0)
    m["clearcoat_roughness"] = kw.get("clearcoat_roughness", 0.01)
    m["specular_color"] = kw.get("specular_color", (0, 0, 0))
    m["use_specular_workflow"] = int(kw.get("use_specular_workflow", 0))
    m["hair_absorption"] = kw.get("hair_absorption", (0, 0, 0))
    m["hair_roughness_lon"] = kw.get("hair_roughness_lon", 0.3)
    m["hair_roughness_azi"] = kw.get("hair_roughness_azi", 0.3)
    m["hair_cuticle_angle"] = kw.get("hair_cuticle_angle", 0.035)
    return m


UPS_CASES = [dict(base_color=(0.8, 0.5, 0.3), roughness=r, metallic=mt) for r in (0.1, 0.5, 1.0) for mt in (0.0, 1.0)] + [
    dict(base_color=(0.2, 0.6, 0.9), roughness=0.4, metallic=0.0, clearcoat=0.8, clearcoat_roughness=0.2),
    dict(base_color=(0.7, 0.7, 0.2), roughness=0.6, use_specular_workflow=1, specular_color=(0.3, 0.5, 0.9)),
]
HAIR_CASES = [
    dict(hair_absorption=(0.3, 0.5, 1.2), hair_roughness_lon=0.3, hair_roughness_azi=0.3, hair_cuticle_angle=0.035, ior=1.55),
    dict(hair_absorption=(0.06, 0.1, 0.2), hair_roughness_lon=0.5, hair_roughness_azi=0.6, hair_cuticle_angle=0.05, ior=1.55),
]
WHITE_HAIR = dict(hair_absorption=(0, 0, 0), hair_roughness_lon=0.4, hair_roughness_azi=0.5, hair_cuticle_angle=0.04, ior=1.55)

# local frames: surfaces n = ng = +z, tangent +x; fibres tangent +x, normal +y
N_SURF, T_SURF = (0.0, 0.0, 1.0), (1.0, 0.0, 0.0)
N_HAIR, T_HAIR = (0.0, 1.0, 0.0), (1.0, 0.0, 0.0)


def pack(n, ng, t, k1, xi, k2):
    cols = [np.atleast_2d(np.asarray(a, dtype=np.float32)) for a in (n, ng, t, k1, xi, k2)]
    count = max(len(c) for c in cols)
    return np.ascontiguousarray(np.concatenate([np.broadcast_to(c, (count, c.shape[1])) for c in cols], axis=1), dtype=np.float32)


def unit(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def ref_model(model, kw):
    if model == _abi.SB_MATERIAL_DIFFUSE:
        return R.Lambert(kw.get("base_color", (1, 1, 1)))
    if model == _abi.SB_MATERIAL_USD_PREVIEW_SURFACE:
        return R.PreviewSurface(kw.get("base_color", (1, 1, 1)), kw.get("roughness", 0.5), kw.get("metallic", 0.0), kw.get("ior", 1.5),
                                kw.get("clearcoat", 0.0), kw.get("clearcoat_roughness", 0.01), kw.get("specular_color", (0, 0, 0)),
                                bool(kw.get("use_specular_workflow", 0)))
    return R.Hair(kw["hair_absorption"], kw["hair_roughness_lon"], kw["hair_roughness_azi"], kw["hair_cuticle_angle"], kw.get("ior", 1.55))


def hair_local(w):
    """world (x = tangent, y = normal, z = binormal) is already the fibre frame"""
    return np.asarray(w, dtype=np.float64)


def warped_grid(pole, n_t=1536, n_phi=1024):
    """Whole sphere about `pole` with cos(theta) = 1 - t^2, t in (0, sqrt 2): cells shrink towards the pole, where a
    sharp lobe centred on it lives.  Returns (dirs (N, 3) float64, weights (N,))."""
    pole = unit(pole)
    a = np.array([1.0, 0, 0]) if abs(pole[0]) < 0.9 else np.array([0, 1.0, 0])
    e1 = unit(np.cross(pole, a))
    e2 = np.cross(pole, e1)
    dt = np.sqrt(2.0) / n_t
    t = (np.arange(n_t) + 0.5) * dt
    ph = (np.arange(n_phi) + 0.5) * 2 * PI / n_phi
    t, ph = np.meshgrid(t, ph, indexing="ij")
    mu = 1.0 - t * t
    st = np.sqrt(np.maximum(1 - mu * mu, 0))
    d = (st * np.cos(ph))[..., None] * e1 + (st * np.sin(ph))[..., None] * e2 + mu[..., None] * pole
    w = 2.0 * t * dt * (2 * PI / n_phi)
    return d.reshape(-1, 3), w.reshape(-1)


# ---- the checks --------------------------------------------------------------------------------------------------
def check_values_against_reference(run, model, kw, rtol=3e-4, seed=1, n=4000):
    """evaluate() == the numpy restatement of the published formulas, for random direction pairs"""
    rng = np.random.default_rng(seed)
    m = material(model, **kw)
    ref = ref_model(model, kw)
    hair = model == _abi.SB_MATERIAL_HAIR
    worst = 0.0
    for _ in range(4):
        k1 = unit(rng.normal(size=3))
        if not hair:
            k1[2] = abs(k1[2]) * 0.95 + 0.05
            k1 = unit(k1)
        k2 = unit(rng.normal(size=(n, 3)))
        k1f, k2f = k1.astype(np.float32), k2.astype(np.float32)  # what the implementation sees
        _, ev = run(m, pack(N_HAIR if hair else N_SURF, N_HAIR if hair else N_SURF, T_HAIR if hair else T_SURF, k1f, (0.5, 0.5, 0.5, 0.5), k2f))
        if hair:
            f, pdf = ref.eval(k1f.astype(np.float64), k2f.astype(np.float64))
            want = np.concatenate([np.zeros_like(f), f, pdf[:, None]], axis=1)
        else:
            d, g, pdf = ref.eval(k1f.astype(np.float64), k2f.astype(np.float64))
            want = np.concatenate([d, g, pdf[:, None]], axis=1)
        scale = np.abs(want).max(axis=0) + 1e-12
        err = np.abs(ev - want) / (np.abs(want) + 1e-3 * scale)
        worst = max(worst, float(err.max()))
    assert worst <= rtol, f"evaluate() deviates from the published formulas by {worst:.2e} (relative)"


def check_sample_evaluate_consistency(run, model, kw, seed=2, n=50000):
    """sample().pdf == evaluate(k2).pdf and bsdf_over_pdf * pdf == f * cos, event bits sane"""
    rng = np.random.default_rng(seed)
    m = material(model, **kw)
    hair = model == _abi.SB_MATERIAL_HAIR
    k1 = unit(rng.normal(size=(n, 3)))
    if not hair:
        k1[:, 2] = np.abs(k1[:, 2]) * 0.98 + 0.02
        k1 = unit(k1)
    xi = rng.random((n, 4))
    fr = (N_HAIR, N_HAIR, T_HAIR) if hair else (N_SURF, N_SURF, T_SURF)
    s, _ = run(m, pack(*fr, k1, xi, (0, 0, 1)))
    ok = s[:, 7] != EV_ABSORB
    assert ok.mean() > (0.99 if hair else 0.5)
    _, e = run(m, pack(*fr, k1, xi, s[:, :3]))
    f = e[:, :3] + e[:, 3:6]
    np.testing.assert_allclose(np.linalg.norm(s[ok, :3], axis=1), 1.0, atol=2e-6)
    np.testing.assert_allclose(s[ok, 6], e[ok, 6], rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(s[ok, 3:6] * s[ok, 6:7], f[ok], rtol=3e-4, atol=1e-6)
    ev = s[ok, 7].astype(int)
    assert np.all((ev & (EV_DIFFUSE | EV_GLOSSY)) != 0)
    assert np.all(((ev & EV_REFLECTION) != 0) ^ ((ev & EV_TRANSMISSION) != 0))
    if hair:
        side = (s[ok, :3] * np.asarray(N_HAIR)).sum(1)
        assert np.all(((ev & EV_TRANSMISSION) != 0) == (side < 0))
        assert ((ev & EV_TRANSMISSION) != 0).mean() > 0.05  # the TT lobe is live: `inside` toggles on the path
    else:
        assert np.all((ev & EV_TRANSMISSION) == 0) and np.all(s[ok, 2] > 0)


def check_pdf_normalisation_and_energy(run, model, kw, k1, seed=3, n_mc=400000, white=False):
    """quadrature of evaluate().pdf over the sphere + absorbed fraction of sample() == 1; quadrature of f cos ==
    the Monte-Carlo mean of bsdf_over_pdf (the one-bounce white furnace); <= 1 for white materials"""
    rng = np.random.default_rng(seed)
    m = material(model, **kw)
    hair = model == _abi.SB_MATERIAL_HAIR
    k1 = unit(k1)
    fr = (N_HAIR, N_HAIR, T_HAIR) if hair else (N_SURF, N_SURF, T_SURF)
    if hair:
        dirs, w = R.sphere_grid(1024, 2048)
        w = np.full(len(dirs), w)
    else:
        mirror = np.array([-k1[0], -k1[1], k1[2]])
        dirs, w = warped_grid(mirror)
    _, e = run(m, pack(*fr, k1, (0.5, 0.5, 0.5, 0.5), dirs))
    int_pdf = float((e[:, 6].astype(np.float64) * w).sum())
    albedo = ((e[:, :3] + e[:, 3:6]).astype(np.float64) * w[:, None]).sum(0)
    s, _ = run(m, pack(*fr, k1, rng.random((n_mc, 4)), (0, 0, 1)))
    absorbed = float((s[:, 7] == EV_ABSORB).mean())
    assert abs(int_pdf + absorbed - 1.0) <= 2e-3, f"pdf integrates to {int_pdf:.5f}, absorbed fraction {absorbed:.5f}"
    mc = s[:, 3:6].astype(np.float64).mean(0)  # absorbed samples carry weight 0
    sd = s[:, 3:6].astype(np.float64).std(0) / np.sqrt(n_mc)
    assert np.all(np.abs(mc - albedo) <= 5 * sd + 2e-3), f"E[bsdf_over_pdf] {mc} vs quadrature of f cos {albedo}"
    assert np.all(albedo <= 1.0 + 2e-3), f"directional albedo {albedo} exceeds 1"
    if white:
        assert np.all(np.abs(albedo - 1.0) <= 2e-3), f"white furnace: albedo {albedo} instead of 1"
    return albedo


def check_sampling_matches_pdf(run, model, kw, k1, seed=4, n=600000):
    """the directions sample() produces are distributed like evaluate().pdf: coarse histogram vs binned quadrature"""
    rng = np.random.default_rng(seed)
    m = material(model, **kw)
    hair = model == _abi.SB_MATERIAL_HAIR
    k1 = unit(k1)
    fr = (N_HAIR, N_HAIR, T_HAIR) if hair else (N_SURF, N_SURF, T_SURF)
    nb_mu, nb_ph, sub = 16, 32, 24
    dirs, w = R.sphere_grid(nb_mu * sub, nb_ph * sub)
    _, e = run(m, pack(*fr, k1, (0.5, 0.5, 0.5, 0.5), dirs))
    p = (e[:, 6].astype(np.float64) * w).reshape(nb_mu, sub, nb_ph, sub).sum(axis=(1, 3))
    s, _ = run(m, pack(*fr, k1, rng.random((n, 4)), (0, 0, 1)))
    ok = s[:, 7] != EV_ABSORB
    d = s[ok, :3].astype(np.float64)
    imu = np.clip(((d[:, 2] + 1.0) * 0.5 * nb_mu).astype(int), 0, nb_mu - 1)
    iph = np.clip((np.mod(np.arctan2(d[:, 1], d[:, 0]), 2 * PI) / (2 * PI) * nb_ph).astype(int), 0, nb_ph - 1)
    hist = np.zeros((nb_mu, nb_ph))
    np.add.at(hist, (imu, iph), 1.0)
    freq = hist / n
    sigma = np.sqrt(np.maximum(p, 1e-12) / n)
    bad = np.abs(freq - p) > 6 * sigma + 2e-3 * p + 2e-5
    assert not bad.any(), f"{bad.sum()} of {bad.size} histogram bins disagree with the pdf (worst {np.abs(freq - p).max():.2e})"


def check_reciprocity(run, model, kw, seed=5, n=5000):
    rng = np.random.default_rng(seed)
    m = material(model, **kw)
    a = unit(rng.normal(size=(n, 3)))
    b = unit(rng.normal(size=(n, 3)))
    a[:, 2] = np.abs(a[:, 2]) * 0.9 + 0.1
    b[:, 2] = np.abs(b[:, 2]) * 0.9 + 0.1
    a, b = unit(a), unit(b)
    _, e1 = run(m, pack(N_SURF, N_SURF, T_SURF, a, (0.5,) * 4, b))
    _, e2 = run(m, pack(N_SURF, N_SURF, T_SURF, b, (0.5,) * 4, a))
    f1 = (e1[:, :3] + e1[:, 3:6]) / b[:, 2:3]
    f2 = (e2[:, :3] + e2[:, 3:6]) / a[:, 2:3]
    np.testing.assert_allclose(f1, f2, rtol=5e-4, atol=1e-6)
