"""Geometry helpers shared by the scene generators."""
from __future__ import annotations

import numpy as np

from ..scene import make_vertices

_F = np.float32


def tangent_for(n: np.ndarray) -> np.ndarray:
    """Tangent as the Hydra delegate derives it: normalised cross(n, X or Y) (Mesh.cpp:149-163)."""
    n = np.asarray(n, dtype=np.float64)
    x = np.array([1.0, 0.0, 0.0])
    y = np.array([0.0, 1.0, 0.0])
    t = np.cross(n, x)
    small = np.linalg.norm(t, axis=-1, keepdims=True) < 1e-6
    t = np.where(small, np.cross(n, y), t)
    return t / np.linalg.norm(t, axis=-1, keepdims=True)


def soup(tris: np.ndarray, normals: np.ndarray | None = None):
    """Triangle soup (T,3,3) -> (vertex buffer, index buffer) un-indexed like Mesh.cpp:140-178.
    `normals` (T,3,3) per-corner normals; flat face normals when None."""
    tris = np.asarray(tris, dtype=np.float64).reshape(-1, 3, 3)
    if normals is None:
        fn = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0])
        fn /= np.linalg.norm(fn, axis=-1, keepdims=True)
        normals = np.repeat(fn[:, None, :], 3, axis=1)
    normals = np.asarray(normals, dtype=np.float64).reshape(-1, 3)
    pos = tris.reshape(-1, 3)
    vb = make_vertices(pos.astype(_F), normals=normals.astype(_F), tangents=tangent_for(normals).astype(_F))
    ib = np.arange(len(pos), dtype=np.uint32)
    return vb, ib


def make_quad_mesh(p0, p1, p2, p3):
    """Quad with corners counter-clockwise seen from its front (normal) side."""
    p = [np.asarray(x, dtype=np.float64) for x in (p0, p1, p2, p3)]
    return soup(np.array([[p[0], p[1], p[2]], [p[0], p[2], p[3]]]))


def make_box_mesh(size, rot_y_deg: float = 0.0, center=(0.0, 0.0, 0.0)):
    """Axis-aligned box of `size` (sx,sy,sz) rotated about +Y, outward-facing normals, 12 triangles."""
    sx, sy, sz = [s * 0.5 for s in size]
    c = np.array(
        [[-sx, -sy, -sz], [sx, -sy, -sz], [sx, sy, -sz], [-sx, sy, -sz], [-sx, -sy, sz], [sx, -sy, sz], [sx, sy, sz], [-sx, sy, sz]]
    )
    a = np.radians(rot_y_deg)
    rot = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    c = c @ rot.T + np.asarray(center, dtype=np.float64)
    faces = [(4, 5, 6, 7), (1, 0, 3, 2), (5, 1, 2, 6), (0, 4, 7, 3), (7, 6, 2, 3), (0, 1, 5, 4)]
    tris = []
    for f in faces:
        tris.append([c[f[0]], c[f[1]], c[f[2]]])
        tris.append([c[f[0]], c[f[2]], c[f[3]]])
    return soup(np.array(tris))


def make_icosphere(subdiv: int, smooth: bool = True):
    """Unit icosphere as triangle soup: returns (tris (T,3,3), normals (T,3,3)); 20*4^subdiv triangles."""
    t = (1.0 + 5.0**0.5) / 2.0
    v = np.array(
        [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]],
        dtype=np.float64,
    )
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array(
        [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
         [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]]
    )
    tris = v[f]
    for _ in range(subdiv):
        a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
        ab, bc, ca = a + b, b + c, c + a
        ab /= np.linalg.norm(ab, axis=1, keepdims=True)
        bc /= np.linalg.norm(bc, axis=1, keepdims=True)
        ca /= np.linalg.norm(ca, axis=1, keepdims=True)
        tris = np.concatenate(
            [np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1), np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)]
        )
    normals = tris.copy() if smooth else None
    return tris, normals
