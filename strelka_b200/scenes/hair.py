"""C4: synthetic hair -- cubic B-spline strands on a scalp sphere.

62 500 strands x 16 segments = 1.0 M segments at full size; 19 control points per strand after the
phantom end points that HdStrelkaBasisCurves::_ConvertCurve adds (BasisCurves.cpp:189-232); roots on the
upper hemisphere of an r = 0.1 m sphere (Fibonacci lattice); growth = normal + gravity droop + jitter;
strand length 0.25 m; widths linear 80 um -> 20 um (stored as radii = width/2 like the delegate does);
two rect lights (key 200, fill 50); scalp sphere diffuse 0.2; hair colour (0.8, 0.7, 0.45).
"""
from __future__ import annotations

import numpy as np

from .. import _abi
from ..scene import Scene, UniformLightDesc, light_look_at
from ..settings import default_settings
from .common import make_icosphere, soup


def convert_curve(points: np.ndarray, widths: np.ndarray):
    """_ConvertCurve (BasisCurves.cpp:189-232) for ONE strand: phantom points at both ends, widths -> radii."""
    p0 = points[0] + (points[0] - points[1])
    pn = points[-1] + (points[-1] - points[-2])
    pts = np.concatenate([p0[None], points, pn[None]])
    r = widths * 0.5
    rad = np.concatenate([r[:1], r, r[-1:]])
    return pts.astype(np.float32), rad.astype(np.float32)


def make_hair(width: int = 1024, height: int = 1024, spp_total: int = 1024, depth: int = 6, n_strands: int = 62500,
              segments: int = 16, seed: int = 0x5EED + 4, single_prim: bool = False, material: str = "diffuse"):
    rng = np.random.default_rng(seed)
    s = Scene()
    s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=(1.0, 1.0, 1.0))
    scalp_mat = s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=(0.2, 0.2, 0.2))
    model = _abi.SB_MATERIAL_HAIR if material == "hair" else _abi.SB_MATERIAL_DIFFUSE
    hair_mat = s.addMaterial(model=model, base_color=(0.8, 0.7, 0.45), hair_absorption=(0.3, 0.5, 1.2),
                             hair_roughness_lon=0.3, hair_roughness_azi=0.3, hair_cuticle_angle=0.035)
    tris, nrm = make_icosphere(4)
    vb, ib = soup(tris * 0.1, nrm)
    s.createInstance(_abi.SB_INSTANCE_MESH, s.createMesh(vb, ib), scalp_mat, np.eye(4))
    # roots: Fibonacci lattice on the upper hemisphere
    n_user = segments + 1  # user control points per strand; +2 phantom = segments + 3 -> `segments` segments
    i = np.arange(n_strands) + 0.5
    z = i / n_strands  # cos(theta) in (0,1): upper hemisphere
    phi = i * np.pi * (3.0 - np.sqrt(5.0))
    rxy = np.sqrt(1.0 - z * z)
    normal = np.stack([rxy * np.cos(phi), z, rxy * np.sin(phi)], axis=1)  # +Y up
    root = normal * 0.1
    step = 0.25 / segments
    pts = np.zeros((n_strands, n_user, 3))
    pts[:, 0] = root
    d = normal.copy()
    gravity = np.array([0.0, -1.0, 0.0])
    for k in range(1, n_user):
        d = d + gravity * 0.12 + rng.uniform(-1.0, 1.0, (n_strands, 3)) * 0.04
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        pts[:, k] = pts[:, k - 1] + d * step + rng.uniform(-1.0, 1.0, (n_strands, 3)) * 0.0005
    widths = np.linspace(80e-6, 20e-6, n_user)
    if single_prim:
        # quirk Q14: phantom points once per PRIM while vertexCounts stay per strand
        allp = pts.reshape(-1, 3)
        allw = np.tile(widths, n_strands)
        p, r = convert_curve(allp, allw)
        c = s.createCurve(np.full(n_strands, n_user, dtype=np.uint32), p, r)
        s.createInstance(_abi.SB_INSTANCE_CURVE, c, hair_mat, np.eye(4))
    else:
        # vectorised per-strand conversion, one prim + one instance per strand (how Hydra hands them over)
        p0 = pts[:, :1] + (pts[:, :1] - pts[:, 1:2])
        pn = pts[:, -1:] + (pts[:, -1:] - pts[:, -2:-1])
        allp = np.concatenate([p0, pts, pn], axis=1).astype(np.float32)  # (n, n_user+2, 3)
        r = widths * 0.5
        rad = np.concatenate([r[:1], r, r[-1:]]).astype(np.float32)
        ncp = n_user + 2
        # bulk-append instead of 62 500 createCurve calls (same resulting arrays)
        s.journal = None  # appended behind the API's back
        base_counts, base_points, base_w = s._nccounts, s._ncpoints, s._ncwidths
        s._cpoints.append(allp.reshape(-1, 3))
        s._cwidths.append(np.tile(rad, n_strands))
        s._ccounts.append(np.full(n_strands, ncp, dtype=np.uint32))
        first_curve = len(s.curves)
        for k in range(n_strands):
            s.curves.append((base_counts + k, 1, base_points + k * ncp, ncp, base_w + k * ncp, ncp))
        s._ncpoints += n_strands * ncp
        s._ncwidths += n_strands * ncp
        s._nccounts += n_strands
        ident = np.eye(4, dtype=np.float32).T.reshape(16)
        for k in range(n_strands):
            s.instances.append((ident, _abi.SB_INSTANCE_CURVE, first_curve + k, hair_mat, 0xFFFFFFFF))
    key = light_look_at((0.5, 0.55, 0.5), (0.0, 0.0, 0.0))
    s.createLight(UniformLightDesc(type=0, xform=key, color=(1.0, 1.0, 1.0), intensity=200.0, width=0.5, height=0.5))
    fill = light_look_at((-0.6, 0.2, 0.4), (0.0, 0.0, 0.0))
    s.createLight(UniformLightDesc(type=0, xform=fill, color=(1.0, 1.0, 1.0), intensity=50.0, width=0.5, height=0.5))
    cam = s.getCamera(0)
    cam.setFov(35.0)
    cam.look_at((0.0, 0.05, 0.9), (0.0, 0.0, 0.0))
    st = default_settings(spp_total=spp_total, spp=1)
    st.setAs("render/pt/depth", depth)
    st.setAs("render/pt/tonemapperType", 0)
    st.setAs("render/post/gamma", 0.0)
    return s, st, (width, height)
