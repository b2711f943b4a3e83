"""Shared helpers of the test-suite: seeded random scenes and ray sets."""
import numpy as np

from strelka_b200 import _abi
from strelka_b200.scene import Scene, UniformLightDesc, rotate_matrix, scale_matrix, translate_matrix
from strelka_b200.scenes.common import make_box_mesh, make_icosphere, soup
from strelka_b200.settings import default_settings


def random_scene(seed=0, n_meshes=6, n_instances=14, lights=("rect", "sphere", "distant", "disc"), usd_materials=True):
    """Instanced icospheres/boxes with random rigid+uniform-scale transforms inside a unit-ish room, several
    light types (incl. the never-sampled disc light, quirk Q7, and the distant light's mesh-0 blocker, Q8)."""
    rng = np.random.default_rng(seed)
    s = Scene()
    s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=(1, 1, 1))
    mats = [0]
    for i in range(5):
        if usd_materials and i % 2 == 1:
            mats.append(s.addMaterial(model=_abi.SB_MATERIAL_USD_PREVIEW_SURFACE, base_color=tuple(rng.uniform(0.1, 0.9, 3)),
                                      roughness=float(rng.uniform(0.15, 0.9)), metallic=float(i % 4 == 1), ior=1.5,
                                      clearcoat=float(rng.uniform(0, 1)) if i == 3 else 0.0, clearcoat_roughness=0.05))
        else:
            mats.append(s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=tuple(rng.uniform(0.1, 0.9, 3))))
    # room: a big box seen from inside (flip winding so normals face inward)
    vb, ib = make_box_mesh((3.0, 3.0, 3.0))
    vb = vb.reshape(-1, 3)[:, ::-1].reshape(-1)  # reverse winding per triangle
    from strelka_b200.scene import pack_normal, unpack_normal
    vb["normal"] = pack_normal(-unpack_normal(vb["normal"]))
    room = s.createMesh(vb, ib)
    s.createInstance(_abi.SB_INSTANCE_MESH, room, mats[1], np.eye(4))
    meshes = []
    for m in range(n_meshes):
        if m % 2 == 0:
            tris, nrm = make_icosphere(1 + (m // 2) % 2)
            vb, ib = soup(tris * 0.25, nrm)
        else:
            vb, ib = make_box_mesh(tuple(rng.uniform(0.15, 0.5, 3)))
        meshes.append(s.createMesh(vb, ib))
    for i in range(n_instances):
        t = translate_matrix(rng.uniform(-1.1, 1.1, 3))
        r = rotate_matrix(rng.normal(size=3), float(rng.uniform(0, 360)))
        sc = float(rng.uniform(0.5, 1.6))
        s.createInstance(_abi.SB_INSTANCE_MESH, meshes[i % len(meshes)], mats[int(rng.integers(0, len(mats)))],
                         t @ r @ scale_matrix(sc, sc, sc))
    for kind in lights:
        if kind == "rect":
            xf = translate_matrix((0.2, 1.45, -0.1)) @ rotate_matrix((1, 0, 0), -90.0) @ rotate_matrix((0, 0, 1), 20.0)
            s.createLight(UniformLightDesc(type=0, xform=xf, color=(1.0, 0.9, 0.8), intensity=30.0, width=0.6, height=0.4))
        elif kind == "sphere":
            s.createLight(UniformLightDesc(type=2, xform=translate_matrix((-0.9, 0.9, 0.8)), color=(0.6, 0.8, 1.0), intensity=25.0, radius=0.12))
        elif kind == "distant":
            xf = rotate_matrix((1, 0, 0), -60.0)
            s.createLight(UniformLightDesc(type=3, xform=xf, color=(1.0, 1.0, 1.0), intensity=2.0, radius=0.05, halfAngle=0.05))
        elif kind == "disc":
            xf = translate_matrix((0.9, 1.4, 0.9)) @ rotate_matrix((1, 0, 0), 90.0)
            s.createLight(UniformLightDesc(type=1, xform=xf, color=(1.0, 0.5, 0.5), intensity=20.0, radius=0.2))
    cam = s.getCamera(0)
    cam.setFov(55.0)
    cam.look_at((0.3, 0.2, 1.4), (-0.1, -0.1, -0.3))
    st = default_settings(spp_total=64, spp=1)
    st.setAs("render/pt/tonemapperType", 0)
    st.setAs("render/post/gamma", 0.0)
    return s, st


def random_rays(n, seed=0, extent=1.3, tmax=1e16):
    rng = np.random.default_rng(seed)
    org = rng.uniform(-extent, extent, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # a share of axis-aligned / zero-component directions (division-by-zero paths of the slab test)
    k = n // 10
    d[:k, 0] = 0.0
    d[:k // 2, 1] = 0.0
    nz = np.linalg.norm(d, axis=1, keepdims=True)
    d = np.where(nz > 0, d / np.maximum(nz, 1e-30), np.array([[0.0, 0.0, 1.0]]))
    tm = np.full((n, 1), tmax)
    return np.concatenate([org, np.zeros((n, 1)), d, tm], axis=1).astype(np.float32)
