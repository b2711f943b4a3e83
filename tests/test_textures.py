"""Textured UsdPreviewSurface on the CPU: the oracle's texture model against an independent numpy
statement of the CUDA linear-filter definition, the host emulation of the device shade code against the oracle on a
textured scene, wrap addressing and the v flip."""
import numpy as np

import pyemul
from conftest import rel_rmse
from oracle import pyoracle
from strelka_b200.scenes import make_kitchen
from strelka_b200.scenes.kitchen import procedural_textures


def numpy_bilinear(tex, uv):
    """CUDA Programming Guide, linear filtering with wrap addressing and 1.8 fixed-point weights (float64)"""
    h, w, _ = tex.shape
    t = tex.astype(np.float64) / 255.0
    x = (uv[:, 0] - np.floor(uv[:, 0])) * w - 0.5
    y = (uv[:, 1] - np.floor(uv[:, 1])) * h - 0.5
    i, j = np.floor(x).astype(int), np.floor(y).astype(int)
    a = np.round((x - i) * 256.0) / 256.0
    b = np.round((y - j) * 256.0) / 256.0
    i0, i1, j0, j1 = i % w, (i + 1) % w, j % h, (j + 1) % h
    a, b = a[:, None], b[:, None]
    return (1 - a) * (1 - b) * t[j0, i0] + a * (1 - b) * t[j0, i1] + (1 - a) * b * t[j1, i0] + a * b * t[j1, i1]


def test_oracle_texture_lookup_follows_the_cuda_filter_definition():
    s, _, _ = make_kitchen(32, 18, 1, n_props=4, subdiv=1, textured=True)
    o = pyoracle.OracleScene(s)
    rng = np.random.default_rng(2)
    uv = rng.uniform(-3.0, 4.0, (20000, 2)).astype(np.float32)  # beyond [0, 1): wrap addressing
    for idx in (0, 1):
        got = o.texture_lookup(idx, uv)
        want = numpy_bilinear(s.textures[idx], uv.astype(np.float64))
        np.testing.assert_allclose(got, want, atol=2e-7)
    # texel centres return the texel itself
    h, w, _ = s.textures[0].shape
    centres = np.array([[(3 + 0.5) / w, (5 + 0.5) / h], [(w - 1 + 0.5) / w, 0.5 / h]], dtype=np.float32)
    got = o.texture_lookup(0, centres)
    np.testing.assert_allclose(got[0], s.textures[0][5, 3] / 255.0, atol=1e-6)
    np.testing.assert_allclose(got[1], s.textures[0][0, w - 1] / 255.0, atol=1e-6)


def test_textured_scene_emulated_shade_code_matches_oracle():
    s, st, _ = make_kitchen(64, 36, 4, n_props=30, subdiv=2, textured=True)
    assert any(m["diffuse_texture"] for m in s.materials) and any(m["normal_texture"] for m in s.materials)
    o, e = pyoracle.OracleScene(s), pyemul.EmulScene(s)
    img_o, _, _, _ = o.render(st, 64, 36, 4)
    img_e, _, _ = e.render(st, 64, 36, 4, chunk_max=2)
    assert rel_rmse(img_e, img_o) < 1e-4
    # the textures matter: the untextured twin of the scene renders differently
    s0, st0, _ = make_kitchen(64, 36, 4, n_props=30, subdiv=2, textured=False)
    img_0, _, _, _ = pyoracle.OracleScene(s0).render(st0, 64, 36, 4)
    assert rel_rmse(img_0, img_o) > 0.05
    # debug view 1 shows the normal-mapped shading normal (state.normal after mdlcode_init)
    st.setAs("render/pt/debug", 1)
    n_o, _, _, _ = o.render(st, 64, 36, 1)
    n_e, _, _ = e.render(st, 64, 36, 1, chunk_max=1)
    # (a filter weight that falls on a 1/256 rounding boundary may differ by one step between two implementations)
    assert np.abs(n_e[..., :3] - n_o[..., :3]).max() < 5e-4


def test_procedural_textures_are_seeded():
    a = procedural_textures(np.random.default_rng(1))
    b = procedural_textures(np.random.default_rng(1))
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and a[0].shape == (256, 256, 4) and a[0].dtype == np.uint8
