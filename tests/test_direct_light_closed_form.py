"""Integrator-level pin that does not share code with the oracle or the kernels: one rect light over
a Lambert plane.  The expectation of the reference's estimator -- NEE with its double cosine (Q4), balance-heuristic
weights, cos-scaled emitter hits (Q5), the area-formula light pdf on the emitter side whatever the sampling method --
is integrated over the light by host quadrature (tests/direct_light_ref.py) and compared with converged renders, for
both rectLightSamplingMethods and for depth 1 (NEE only) and depth 2 (NEE + BSDF-sampled emitter hits)."""
import numpy as np
import pytest

import direct_light_ref as D
import direct_light_scene as DS
from oracle import pyoracle

SPP = 1024


def expected_image(scene, method, depth):
    pts = DS.pixel_points(scene)
    light = scene.arrays()["lights"][0]
    img = np.zeros((DS.H, DS.W, 3))
    for y in range(DS.H):
        for x in range(DS.W):
            img[y, x] = D.expected_radiance(light, pts[y, x], (0.0, 1.0, 0.0), DS.RHO, method, depth, n_quad=192)
    return img


def compare(img, want):
    """image mean within 0.5 % (MC error of 24*24*1024 samples is ~0.1 %), every pixel within its own MC error"""
    assert want.min() > 0
    assert abs(img.mean() / want.mean() - 1.0) < 5e-3, (img.mean(), want.mean())
    rel = np.abs(img / want - 1.0)
    assert np.percentile(rel, 99) < 0.08 and rel.max() < 0.15, (np.percentile(rel, 99), rel.max())


@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("depth", [1, 2])
def test_oracle_converges_to_the_closed_form(method, depth):
    scene, st = DS.make(method, depth, SPP)
    img, _, _, _ = pyoracle.OracleScene(scene).render(st, DS.W, DS.H, SPP)
    compare(img[..., :3].astype(np.float64), expected_image(scene, method, depth))


def test_quadrature_is_self_consistent():
    """depth-2 expectation with uniform sampling (weights sum to one pointwise) splits into the two strategies' integrands"""
    scene, _ = DS.make(0, 2, 1)
    light = scene.arrays()["lights"][0]
    x, n = np.array([0.3, 0.0, 0.4]), (0.0, 1.0, 0.0)
    a = D.expected_radiance(light, x, n, DS.RHO, 0, 2, n_quad=256)
    b = D.expected_radiance(light, x, n, DS.RHO, 0, 2, n_quad=512)
    np.testing.assert_allclose(a, b, rtol=1e-5)
    S = D.spherical_rect_solid_angle(*(lambda P: (P[0], P[1] - P[0], P[3] - P[0]))(np.asarray(light["points"], dtype=np.float64)[:, :3]), x)
    assert 0.0 < S < 2 * np.pi


@pytest.mark.gpu
@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("depth", [1, 2])
def test_device_converges_to_the_closed_form(gpu_render, method, depth):
    from strelka_b200 import BufferDesc, BufferFormat, SharedContext

    spp = 4096
    scene, st = DS.make(method, depth, spp)
    r = gpu_render
    r.setScene(scene)
    r.setSharedContext(SharedContext(mSettingsManager=st))
    r._last_settings = None
    r.reset_accumulation()
    buf = r.createBuffer(BufferDesc(DS.W, DS.H, BufferFormat.FLOAT4))
    r.render_iterations(buf, spp)
    img = buf.map().copy()
    buf.destroy()
    want = expected_image(scene, method, depth)
    assert abs(img[..., :3].mean() / want.mean() - 1.0) < 3e-3
    rel = np.abs(img[..., :3] / want - 1.0)
    assert np.percentile(rel, 99) < 0.04 and rel.max() < 0.08
