"""C3: synthetic Kitchen-Set-scale scene: a 6 x 3 x 5 m room + 400 props (icospheres,
subdivision 4 = 5120 triangles each, ~2.05 M triangles), 50 UsdPreviewSurface materials
(diffuseColor U(0.05,0.9)^3, roughness U(0.1,1), metallic 0 (80 %) / 1, ior 1.5), one 1 x 1 m ceiling
rect light (intensity 60).  Every prop is its own mesh copy + instance, like the Hydra delegate's
per-instance baking (RenderPass.cpp:252-257)."""
from __future__ import annotations

import numpy as np

from .. import _abi
from ..scene import Scene, UniformLightDesc, pack_normal, rotate_matrix, translate_matrix, unpack_normal
from ..settings import default_settings
from .common import make_box_mesh, make_icosphere, soup


def procedural_textures(rng, size: int = 256):
    """Two seeded RGBA8 images: a colour checker with soft blobs (diffuseColor input) and a tangent-space normal map of
    bumps (normal input), the kind of UsdUVTexture pair a Kitchen-class asset carries."""
    y, x = np.mgrid[0:size, 0:size] / float(size)
    chk = ((np.floor(x * 8) + np.floor(y * 8)) % 2)[..., None]
    a, b = rng.uniform(0.1, 0.9, 3), rng.uniform(0.1, 0.9, 3)
    col = chk * a + (1 - chk) * b
    col = col * (0.75 + 0.25 * np.sin(2 * np.pi * (3 * x + 2 * y)))[..., None]
    diffuse = np.concatenate([np.clip(col, 0, 1), np.ones((size, size, 1))], axis=2)
    hgt = 0.5 + 0.5 * np.sin(2 * np.pi * 6 * x) * np.sin(2 * np.pi * 6 * y)
    gx, gy = np.gradient(hgt)
    n = np.stack([-gx * 12.0, -gy * 12.0, np.ones_like(hgt)], axis=2)
    n /= np.linalg.norm(n, axis=2, keepdims=True)
    normal = np.concatenate([n * 0.5 + 0.5, np.ones((size, size, 1))], axis=2)
    to8 = lambda im: np.clip(im * 255.0 + 0.5, 0, 255).astype(np.uint8)  # noqa: E731
    return to8(diffuse), to8(normal)


def sphere_uv(p):
    """lat-long st of unit directions, v flipped like the delegate does (RenderPass.cpp:109-114)"""
    u = np.arctan2(p[..., 2], p[..., 0]) / (2 * np.pi) + 0.5
    v = np.arccos(np.clip(p[..., 1], -1, 1)) / np.pi
    return np.stack([u * 2.0, 1.0 - v], axis=-1)  # u repeats twice around: exercises wrap addressing


def make_kitchen(width: int = 1920, height: int = 1080, spp_total: int = 2048, depth: int = 4, n_props: int = 400, subdiv: int = 4,
                 n_materials: int = 50, seed: int = 0x5EED + 3, textured: bool = False):
    rng = np.random.default_rng(seed)
    s = Scene()
    s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=(1.0, 1.0, 1.0))
    mats = []
    for _ in range(n_materials):
        mats.append(s.addMaterial(model=_abi.SB_MATERIAL_USD_PREVIEW_SURFACE, base_color=tuple(rng.uniform(0.05, 0.9, 3)),
                                  roughness=float(rng.uniform(0.1, 1.0)), metallic=float(rng.uniform() > 0.8), ior=1.5, opacity=1.0))
    if textured:
        # the textured variant: every third material reads diffuseColor from a texture, every fifth
        # has a normal map as well; own generator so that the untextured scene keeps its random stream
        trng = np.random.default_rng(seed + 1000)
        tex = [tuple(s.addTexture(t) for t in procedural_textures(trng)) for _ in range(4)]
        for i in range(0, n_materials, 3):
            d, nm = tex[(i // 3) % len(tex)]
            s.materials[mats[i]]["diffuse_texture"] = d
            if i % 5 == 0:
                s.materials[mats[i]]["normal_texture"] = nm
    wall = s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=(0.7, 0.7, 0.68))
    vb, ib = make_box_mesh((6.0, 3.0, 5.0))
    vb = vb.reshape(-1, 3)[:, ::-1].reshape(-1)  # seen from inside
    vb["normal"] = pack_normal(-unpack_normal(vb["normal"]))
    s.createInstance(_abi.SB_INSTANCE_MESH, s.createMesh(vb, ib), wall, np.eye(4))
    tris, nrm = make_icosphere(subdiv)
    for i in range(n_props):
        r = float(rng.uniform(0.05, 0.3))
        pos = rng.uniform((-2.7, -1.2, -2.2), (2.7, 1.2, 2.2))
        vb, ib = soup(tris * r, nrm)
        if textured:
            from ..scene import pack_uv

            vb["uv"] = pack_uv(sphere_uv(tris.reshape(-1, 3)))
        m = s.createMesh(vb, ib)
        s.createInstance(_abi.SB_INSTANCE_MESH, m, mats[int(rng.integers(0, n_materials))], translate_matrix(pos))
    xf = translate_matrix((0.0, 1.499, 0.0)) @ rotate_matrix((1, 0, 0), -90.0)
    s.createLight(UniformLightDesc(type=0, xform=xf, color=(1.0, 1.0, 1.0), intensity=60.0, width=1.0, height=1.0))
    cam = s.getCamera(0)
    cam.setFov(60.0)
    cam.look_at((0.0, 0.2, 2.45), (0.0, -0.2, 0.0))
    st = default_settings(spp_total=spp_total, spp=1)
    st.setAs("render/pt/depth", depth)
    st.setAs("render/pt/tonemapperType", 0)
    st.setAs("render/post/gamma", 0.0)
    return s, st, (width, height)
