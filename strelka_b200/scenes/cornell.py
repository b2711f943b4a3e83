"""C2 (and the C1 substitute): Cornell-box-style scene, diffuse + one rect light.

Box 1x1x1 m centred at the origin, open towards +Z; white floor/ceiling/back (0.73), red left wall,
green right wall; a short 0.3 m cube rotated -18 deg and a tall 0.3x0.6x0.3 m block rotated +15 deg;
34 un-indexed triangles; one 0.25x0.25 m rect light 1 mm below the ceiling facing -Y,
color (1,1,1) x intensity 40; camera at (0,0,2.2) looking down -Z, fovY 40 deg.
"""
from __future__ import annotations

import numpy as np

from .. import _abi
from ..scene import Scene, UniformLightDesc, rotate_matrix, translate_matrix
from ..settings import default_settings
from .common import make_box_mesh, make_quad_mesh


def make_cornell(width: int = 1024, height: int = 1024, spp_total: int = 256, depth: int = 4, rect_method: int = 0,
                 light_intensity: float = 40.0):
    s = Scene()
    # material slot 0 = default_material injected by the backend's init() (OptixRender.cpp:1091-1097)
    s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=(1.0, 1.0, 1.0))
    white = s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=(0.73, 0.73, 0.73))
    red = s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=(0.65, 0.05, 0.05))
    green = s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=(0.12, 0.45, 0.15))
    h = 0.5
    ident = np.eye(4)
    walls = [
        # corners counter-clockwise seen from inside the room
        (((-h, -h, h), (h, -h, h), (h, -h, -h), (-h, -h, -h)), white),  # floor, normal +Y
        (((-h, h, -h), (h, h, -h), (h, h, h), (-h, h, h)), white),  # ceiling, normal -Y
        (((-h, -h, -h), (h, -h, -h), (h, h, -h), (-h, h, -h)), white),  # back, normal +Z
        (((-h, -h, h), (-h, -h, -h), (-h, h, -h), (-h, h, h)), red),  # left, normal +X
        (((h, -h, -h), (h, -h, h), (h, h, h), (h, h, -h)), green),  # right, normal -X
    ]
    expect = [(0, 1, 0), (0, -1, 0), (0, 0, 1), (1, 0, 0), (-1, 0, 0)]
    for (corners, mat), n in zip(walls, expect):
        vb, ib = make_quad_mesh(*corners)
        fn = np.cross(vb["pos"][1] - vb["pos"][0], vb["pos"][2] - vb["pos"][0])
        assert np.dot(fn, n) > 0, "wall winding must face into the room"
        m = s.createMesh(vb, ib)
        s.createInstance(_abi.SB_INSTANCE_MESH, m, mat, ident)
    # the two blocks are authored at the origin and placed by their instance transform
    vb, ib = make_box_mesh((0.3, 0.3, 0.3))
    m = s.createMesh(vb, ib)
    s.createInstance(_abi.SB_INSTANCE_MESH, m, white, translate_matrix((0.17, -0.35, 0.15)) @ rotate_matrix((0, 1, 0), -18.0))
    vb, ib = make_box_mesh((0.3, 0.6, 0.3))
    m = s.createMesh(vb, ib)
    s.createInstance(_abi.SB_INSTANCE_MESH, m, white, translate_matrix((-0.17, -0.2, -0.15)) @ rotate_matrix((0, 1, 0), 15.0))
    # rect light: local XY quad emitting towards local -Z; rotate so that -Z -> -Y
    xf = translate_matrix((0.0, h - 0.001, 0.0)) @ rotate_matrix((1, 0, 0), -90.0)
    s.createLight(UniformLightDesc(type=0, xform=xf, color=(1.0, 1.0, 1.0), intensity=light_intensity, width=0.25, height=0.25))
    cam = s.getCamera(0)
    cam.setFov(40.0)
    cam.look_at((0.0, 0.0, 2.2), (0.0, 0.0, 0.0))
    settings = default_settings(spp_total=spp_total, spp=1)
    settings.setAs("render/pt/depth", depth)
    settings.setAs("render/pt/rectLightSamplingMethod", rect_method)
    settings.setAs("render/pt/tonemapperType", 0)
    settings.setAs("render/post/gamma", 0.0)
    return s, settings, (width, height)
