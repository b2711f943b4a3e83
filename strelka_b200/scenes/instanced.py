"""C5: synthetic 10 M-triangle instanced scene: 100 prototype meshes (icosphere subdiv 4
displaced by seeded value noise, 5120 triangles each) x 2000 instances on a jittered 20 x 10 x 10 grid,
random rotation, uniform scale U(0.7,1.3) = 10.24 M instanced triangles, 8 default_material colours,
4 rect lights 2 x 2 m (intensity 80) above the grid.

duplicate=True reproduces the Hydra delegate's flattening (one mesh copy per instance, quirk Q16, ~1 GB of
host vertices); duplicate=False hands the backend shared prototypes + instance transforms -- the device
flattens them to the same world-space triangles either way, so the images are identical."""
from __future__ import annotations

import numpy as np

from .. import _abi
from ..scene import Scene, UniformLightDesc, rotate_matrix, scale_matrix, translate_matrix
from ..settings import default_settings
from .common import make_icosphere, soup


def make_instanced(width: int = 3840, height: int = 2160, spp_total: int = 4096, depth: int = 4, n_protos: int = 100,
                   n_instances: int = 2000, subdiv: int = 4, duplicate: bool = False, seed: int = 0x5EED + 5):
    rng = np.random.default_rng(seed)
    s = Scene()
    s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=(1.0, 1.0, 1.0))
    mats = [s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=tuple(rng.uniform(0.15, 0.9, 3))) for _ in range(8)]
    ground = s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=(0.5, 0.5, 0.5))
    tris, _ = make_icosphere(subdiv)
    protos = []
    for p in range(n_protos):
        # smooth seeded displacement: a few random low-frequency lobes on the unit sphere
        k = rng.normal(size=(4, 3))
        amp = rng.uniform(0.05, 0.2, 4)
        v = tris.reshape(-1, 3)
        disp = 1.0 + sum(a * np.sin(3.0 * (v @ kk)) for a, kk in zip(amp, k))
        vd = (v * disp[:, None]).reshape(-1, 3, 3)
        fn = np.cross(vd[:, 1] - vd[:, 0], vd[:, 2] - vd[:, 0])
        fn /= np.linalg.norm(fn, axis=1, keepdims=True)
        protos.append((vd * 0.35, np.repeat(fn[:, None, :], 3, axis=1)))
    proto_mesh = None
    if not duplicate:
        proto_mesh = [s.createMesh(*soup(t, n)) for (t, n) in protos]
    nx, ny, nz = 20, 10, 10
    for i in range(n_instances):
        gx, gy, gz = i % nx, (i // nx) % ny, i // (nx * ny)
        pos = np.array([gx - nx / 2 + 0.5, gy + 0.5, gz - nz / 2 + 0.5]) + rng.uniform(-0.25, 0.25, 3)
        sc = float(rng.uniform(0.7, 1.3))
        xf = translate_matrix(pos) @ rotate_matrix(rng.normal(size=3), float(rng.uniform(0, 360))) @ scale_matrix(sc, sc, sc)
        if duplicate:
            m = s.createMesh(*soup(*protos[i % n_protos]))
        else:
            m = proto_mesh[i % n_protos]
        s.createInstance(_abi.SB_INSTANCE_MESH, m, mats[i % 8], xf)
    from .common import make_quad_mesh
    vb, ib = make_quad_mesh((-14, 0, 9), (14, 0, 9), (14, 0, -9), (-14, 0, -9))
    s.createInstance(_abi.SB_INSTANCE_MESH, s.createMesh(vb, ib), ground, np.eye(4))
    for lx, lz in ((-5, -2.5), (5, -2.5), (-5, 2.5), (5, 2.5)):
        xf = translate_matrix((lx, 13.0, lz)) @ rotate_matrix((1, 0, 0), -90.0)
        s.createLight(UniformLightDesc(type=0, xform=xf, color=(1.0, 1.0, 1.0), intensity=80.0, width=2.0, height=2.0))
    cam = s.getCamera(0)
    cam.setFov(50.0)
    cam.look_at((0.0, 7.0, 19.0), (0.0, 4.5, 0.0))
    st = default_settings(spp_total=spp_total, spp=1)
    st.setAs("render/pt/depth", depth)
    st.setAs("render/pt/tonemapperType", 0)
    st.setAs("render/post/gamma", 0.0)
    return s, st, (width, height)
