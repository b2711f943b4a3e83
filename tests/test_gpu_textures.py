"""Textured UsdPreviewSurface on the device: the texture unit's filtered lookups against the oracle's
software model of the CUDA filter definition, and a textured C3 variant against the oracle (-m gpu)."""
import numpy as np
import pytest

from conftest import rel_rmse
from oracle import pyoracle
from strelka_b200 import BufferDesc, BufferFormat, SharedContext
from strelka_b200.scenes import make_kitchen

pytestmark = pytest.mark.gpu


def test_hardware_lookups_match_the_filter_model(gpu_render):
    s, _, _ = make_kitchen(32, 18, 1, n_props=4, subdiv=1, textured=True)
    gpu_render.setScene(s)
    o = pyoracle.OracleScene(s)
    rng = np.random.default_rng(4)
    uv = rng.uniform(-3.0, 4.0, (200000, 2)).astype(np.float32)
    for idx in range(len(s.textures)):
        hw = gpu_render.test_texture(idx, uv)
        sw = o.texture_lookup(idx, uv)
        d = np.abs(hw - sw)
        # the hardware keeps 8 fractional bits of the filter weights: two implementations may disagree by one weight
        # step (1/256 of the texel-to-texel difference) where the fraction lands on a rounding boundary
        assert d.max() <= 1.5 / 256.0, d.max()
        assert np.percentile(d, 99) <= 2e-3 and d.mean() <= 3e-4, (np.percentile(d, 99), d.mean())
    h, w, _ = s.textures[0].shape
    centres = np.array([[(3 + 0.5) / w, (5 + 0.5) / h], [(w - 1 + 0.5) / w, 0.5 / h]], dtype=np.float32)
    hw = gpu_render.test_texture(0, centres)
    np.testing.assert_allclose(hw[0], s.textures[0][5, 3] / 255.0, atol=1e-6)
    np.testing.assert_allclose(hw[1], s.textures[0][0, w - 1] / 255.0, atol=1e-6)


def _render(r, scene, settings, w, h, n):
    r.setScene(scene)
    r.setSharedContext(SharedContext(mSettingsManager=settings))
    r._last_settings = None
    r.reset_accumulation()
    buf = r.createBuffer(BufferDesc(w, h, BufferFormat.FLOAT4))
    r.render_iterations(buf, n)
    img = buf.map().copy()
    buf.destroy()
    return img


def test_textured_c3_variant_matches_oracle(gpu_render):
    s, st, _ = make_kitchen(192, 108, 8, n_props=60, subdiv=3, textured=True)
    img_g = _render(gpu_render, s, st, 192, 108, 8)
    img_o, _, _, _ = pyoracle.OracleScene(s).render(st, 192, 108, 8)
    assert rel_rmse(img_g, img_o) <= 1e-3
    assert abs(img_g[..., :3].mean() / img_o[..., :3].mean() - 1.0) <= 5e-3
    s0, st0, _ = make_kitchen(192, 108, 8, n_props=60, subdiv=3, textured=False)
    assert rel_rmse(_render(gpu_render, s0, st0, 192, 108, 8), img_o) > 0.05  # the textures matter
    st.setAs("render/pt/debug", 1)  # the normal-mapped shading normal
    n_g = _render(gpu_render, s, st, 192, 108, 1)
    n_o, _, _, _ = pyoracle.OracleScene(s).render(st, 192, 108, 1)
    assert np.abs(n_g[..., :3] - n_o[..., :3]).max() < 2e-3


def test_full_size_textured_c3_windows_match_oracle(gpu_render):
    """the textured variant at the size of BASELINE.json configs[2] (2.05 M triangles, 1920 x 1080, sppTotal 2048)"""
    s, st, (w, h) = make_kitchen(1920, 1080, 2048, textured=True)
    img = _render(gpu_render, s, st, w, h, 1)
    o = pyoracle.OracleScene(s)
    for (x0, y0, ww, wh) in [(832, 412, 256, 256), (1500, 800, 256, 128)]:
        xs, ys = np.meshgrid(np.arange(x0, x0 + ww), np.arange(y0, y0 + wh))
        ref = o.path_radiance(st, w, h, xs.reshape(-1), ys.reshape(-1), np.zeros(xs.size)).reshape(wh, ww, 3)
        got = img[y0:y0 + wh, x0:x0 + ww, :3]
        assert ref.mean() > 0
        # texture filter weights on a 1/256 boundary: isolated texels may differ by one weight step
        assert rel_rmse(got, ref) <= 1e-3
        assert abs(got.mean() / ref.mean() - 1.0) <= 5e-3
