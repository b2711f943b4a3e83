"""One Lambert ground plane under one rect light, seen from above at an angle (the light is outside the frame): the
scene of the closed-form direct-lighting pins (tests/direct_light_ref.py)."""
import numpy as np

from strelka_b200 import _abi
from strelka_b200.scene import Scene, UniformLightDesc, rotate_matrix, translate_matrix
from strelka_b200.scenes.common import make_quad_mesh
from strelka_b200.settings import default_settings

RHO = (0.6, 0.7, 0.8)
W = H = 24


def make(method: int, depth: int, spp_total: int):
    s = Scene()
    s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=(1, 1, 1))
    mat = s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=RHO)
    vb, ib = make_quad_mesh((-6, 0, 6), (6, 0, 6), (6, 0, -6), (-6, 0, -6))  # normal +Y
    s.createInstance(_abi.SB_INSTANCE_MESH, s.createMesh(vb, ib), mat, np.eye(4))
    xf = translate_matrix((0.2, 1.5, -0.3)) @ rotate_matrix((0, 1, 0), 25.0) @ rotate_matrix((1, 0, 0), -90.0)  # emits towards -Y
    s.createLight(UniformLightDesc(type=0, xform=xf, color=(1.0, 0.8, 0.6), intensity=30.0, width=1.0, height=0.6))
    cam = s.getCamera(0)
    cam.setFov(24.0)
    cam.look_at((0.0, 1.6, 3.0), (0.0, 0.0, 0.0))
    st = default_settings(spp_total=spp_total, spp=1)
    st.setAs("render/pt/depth", depth)
    st.setAs("render/pt/rectLightSamplingMethod", method)
    st.setAs("render/pt/tonemapperType", 0)
    st.setAs("render/post/gamma", 0.0)
    return s, st


def pixel_points(scene, w=W, h=H):
    """world points on the plane y = 0 under the pixel centres (row 0 = bottom scanline, quirk Q15)"""
    cam = scene.getCamera(0)
    cam.updateViewMatrix()
    view = np.asarray(cam.view_glm(), dtype=np.float64).reshape(4, 4).T  # glm column-major storage -> matrix
    v2w = np.linalg.inv(view)
    origin = v2w[:3, 3]
    t = np.tan(np.radians(cam.fov) / 2.0)
    pts = np.zeros((h, w, 3))
    for y in range(h):
        for x in range(w):
            ndc = ((x + 0.5) / w * 2 - 1, (y + 0.5) / h * 2 - 1)
            dv = np.array([ndc[0] * t * (w / h), ndc[1] * t, -1.0])
            dw = v2w[:3, :3] @ dv
            dw /= np.linalg.norm(dw)
            k = -origin[1] / dw[1]
            pts[y, x] = origin + k * dw
    return pts
