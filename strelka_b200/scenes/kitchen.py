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


def make_kitchen(width: int = 1920, height: int = 1080, spp_total: int = 2048, depth: int = 4, n_props: int = 400, subdiv: int = 4,
                 n_materials: int = 50, seed: int = 0x5EED + 3):
    rng = np.random.default_rng(seed)
    s = Scene()
    s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=(1.0, 1.0, 1.0))
    mats = []
    for _ in range(n_materials):
        mats.append(s.addMaterial(model=_abi.SB_MATERIAL_USD_PREVIEW_SURFACE, base_color=tuple(rng.uniform(0.05, 0.9, 3)),
                                  roughness=float(rng.uniform(0.1, 1.0)), metallic=float(rng.uniform() > 0.8), ior=1.5, opacity=1.0))
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
