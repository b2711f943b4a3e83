"""oka::Scene mirror (reference: include/scene/scene.h, src/scene/scene.cpp) -- the flat CPU
arrays a render backend consumes, plus the packing helpers of the Hydra delegate
(src/HdStrelka/RenderPass.cpp:53-67) and the light bookkeeping of Scene::createLight/updateLight
(scene.cpp:306-408), restated glm-free with numpy.

Transforms are 4x4 numpy arrays in math convention (p_world = M @ p_object); they are stored in
glm's column-major order when handed to the C ABI (sb_instance.transform).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _abi
from ._abi import (
    CURVE_DTYPE,
    INSTANCE_DTYPE,
    LIGHT_DTYPE,
    MATERIAL_DTYPE,
    MESH_DTYPE,
    VERTEX_DTYPE,
    SB_INSTANCE_CURVE,
    SB_INSTANCE_LIGHT,
    SB_INSTANCE_MESH,
)
from .camera import Camera

_F = np.float32


def pack_normal(n) -> np.ndarray:
    """packNormal, RenderPass.cpp:53-59 / scene.cpp:111-117 (10-10-10 bits), float32 arithmetic."""
    n = np.asarray(n, dtype=_F)
    q = ((n + _F(1.0)) / _F(2.0) * _F(511.99999)).astype(np.uint32)
    return (q[..., 0] + (q[..., 1] << np.uint32(10)) + (q[..., 2] << np.uint32(20))).astype(np.uint32)


def unpack_normal(v) -> np.ndarray:
    """unpackNormal, closest_hit.cu:236-244."""
    v = np.asarray(v, dtype=np.uint32)
    z = ((v & np.uint32(0xFFF00000)) >> np.uint32(20)).astype(_F) / _F(511.99999) * _F(2.0) - _F(1.0)
    y = ((v & np.uint32(0x000FFC00)) >> np.uint32(10)).astype(_F) / _F(511.99999) * _F(2.0) - _F(1.0)
    x = (v & np.uint32(0x000003FF)).astype(_F) / _F(511.99999) * _F(2.0) - _F(1.0)
    return np.stack([x, y, z], axis=-1)


def pack_uv(uv) -> np.ndarray:
    """packUV, RenderPass.cpp:61-67 (16-16 bits over [-10, 10])."""
    uv = np.asarray(uv, dtype=_F)
    q = ((uv + _F(10.0)) / _F(20.0) * _F(16383.99999)).astype(np.uint32)
    return (q[..., 0] + (q[..., 1] << np.uint32(16))).astype(np.uint32)


def make_vertices(pos, normals=None, tangents=None, uvs=None) -> np.ndarray:
    pos = np.asarray(pos, dtype=_F).reshape(-1, 3)
    vb = np.zeros(len(pos), dtype=VERTEX_DTYPE)
    vb["pos"] = pos
    if normals is not None:
        vb["normal"] = pack_normal(np.asarray(normals, dtype=_F).reshape(-1, 3))
    if tangents is not None:
        vb["tangent"] = pack_normal(np.asarray(tangents, dtype=_F).reshape(-1, 3))
    if uvs is not None:
        vb["uv"] = pack_uv(np.asarray(uvs, dtype=_F).reshape(-1, 2))
    else:
        vb["uv"] = pack_uv(np.zeros((len(pos), 2), dtype=_F))
    return vb


def scale_matrix(sx, sy, sz) -> np.ndarray:
    m = np.eye(4)
    m[0, 0], m[1, 1], m[2, 2] = sx, sy, sz
    return m


def translate_matrix(t) -> np.ndarray:
    m = np.eye(4)
    m[:3, 3] = t
    return m


def rotate_matrix(axis, degrees) -> np.ndarray:
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    a = np.radians(degrees)
    c, s = np.cos(a), np.sin(a)
    x, y, z = axis
    m = np.eye(4)
    m[:3, :3] = [
        [c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
        [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
        [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)],
    ]
    return m


def light_look_at(pos, target, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """Transform for a light at `pos` whose emission axis (local -Z, scene.cpp:353-408) points at `target`."""
    pos = np.asarray(pos, dtype=np.float64)
    d = np.asarray(target, dtype=np.float64) - pos
    d /= np.linalg.norm(d)
    z = -d
    x = np.cross(np.asarray(up, dtype=np.float64), z)
    if np.linalg.norm(x) < 1e-6:
        x = np.cross(np.array([1.0, 0.0, 0.0]), z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, pos
    return m


@dataclass
class UniformLightDesc:
    """Scene::UniformLightDesc, scene.h:158-179 (only the useXform path the Hydra delegate uses)."""

    type: int = 0  # 0 rect, 1 disc, 2 sphere, 3 distant
    xform: np.ndarray = field(default_factory=lambda: np.eye(4))
    color: tuple = (1.0, 1.0, 1.0)
    intensity: float = 1.0
    width: float = 1.0
    height: float = 1.0
    radius: float = 1.0
    halfAngle: float = 0.0  # noqa: N815


class Scene:
    """Flat scene arrays, same getters as oka::Scene."""

    def __init__(self):
        self._vb: list[np.ndarray] = []
        self._ib: list[np.ndarray] = []
        self._nverts = 0
        self._nidx = 0
        self.meshes: list[tuple] = []
        self.curves: list[tuple] = []
        self._cpoints: list[np.ndarray] = []
        self._cwidths: list[np.ndarray] = []
        self._ccounts: list[np.ndarray] = []
        self._ncpoints = 0
        self._ncwidths = 0
        self._nccounts = 0
        self.instances: list[tuple] = []
        self.lights: list[np.ndarray] = []
        self.light_descs: list[UniformLightDesc] = []
        self.materials: list[np.ndarray] = []
        self.textures: list[np.ndarray] = []  # (h, w, 4) uint8, row 0 = first row of the image file
        self.cameras: list[Camera] = [Camera()]
        self._rect_mesh = self._disc_mesh = self._sphere_mesh = -1
        self._keep = None
        # journal of the oka::Scene API calls that built this scene, in order (tools replay it through the reference's own
        # scene.cpp: adapter/adapter_ref_test.cpp).  None once an entry was appended behind the API's back.
        self.journal: list | None = []

    # ---- meshes / instances (scene.cpp:15-88) --------------------------------------------------
    def createMesh(self, vb: np.ndarray, ib) -> int:  # noqa: N802
        vb = np.ascontiguousarray(vb, dtype=VERTEX_DTYPE)
        ib = np.ascontiguousarray(ib, dtype=np.uint32).reshape(-1)
        mesh_id = len(self.meshes)
        if self.journal is not None and not getattr(self, "_in_light", False):
            self.journal.append(("mesh", vb, ib))
        self.meshes.append((self._nidx, len(ib), self._nverts, len(vb)))
        self._vb.append(vb)
        self._ib.append(ib)
        self._nverts += len(vb)
        self._nidx += len(ib)
        return mesh_id

    def createInstance(self, type_: int, geom_id: int, material_id: int, transform, light_id: int = 0xFFFFFFFF) -> int:  # noqa: N802
        t = np.asarray(transform, dtype=np.float64).reshape(4, 4)
        if self.journal is not None and not getattr(self, "_in_light", False):
            self.journal.append(("instance", t.T.astype(_F).reshape(16), type_, geom_id, material_id & 0xFFFFFFFF))
        self.instances.append((t.T.astype(_F).reshape(16), type_, geom_id, material_id & 0xFFFFFFFF, light_id & 0xFFFFFFFF))
        return len(self.instances) - 1

    def addMaterial(self, **kw) -> int:  # noqa: N802
        """Scene::addMaterial (scene.cpp:90-96) with the description pre-resolved to sb_material."""
        m = np.zeros((), dtype=MATERIAL_DTYPE)
        m["model"] = kw.get("model", _abi.SB_MATERIAL_DIFFUSE)
        m["base_color"] = kw.get("base_color", (1.0, 1.0, 1.0))
        m["roughness"] = kw.get("roughness", 0.5)
        m["metallic"] = kw.get("metallic", 0.0)
        m["ior"] = kw.get("ior", 1.5)
        m["opacity"] = kw.get("opacity", 1.0)
        m["clearcoat"] = kw.get("clearcoat", 0.0)
        m["clearcoat_roughness"] = kw.get("clearcoat_roughness", 0.01)
        m["specular_color"] = kw.get("specular_color", (0.0, 0.0, 0.0))
        m["use_specular_workflow"] = kw.get("use_specular_workflow", 0)
        m["hair_absorption"] = kw.get("hair_absorption", (0.0, 0.0, 0.0))
        m["hair_roughness_lon"] = kw.get("hair_roughness_lon", 0.3)
        m["hair_roughness_azi"] = kw.get("hair_roughness_azi", 0.3)
        m["hair_cuticle_angle"] = kw.get("hair_cuticle_angle", 0.035)
        m["diffuse_texture"] = kw.get("diffuse_texture", 0)  # 1-based index into self.textures, 0 = none
        m["normal_texture"] = kw.get("normal_texture", 0)
        if self.journal is not None:
            self.journal.append(("material", m))
        self.materials.append(m)
        return len(self.materials) - 1

    def addTexture(self, rgba8) -> int:  # noqa: N802
        """Register an 8-bit RGBA image (what stbi_load(..., STBI_rgb_alpha) hands OptiXRender::loadTextureFromFile,
        OptixRender.cpp:1191-1268).  Returns the 1-based index materials refer to (0 = no texture)."""
        a = np.ascontiguousarray(rgba8, dtype=np.uint8)
        assert a.ndim == 3 and a.shape[2] == 4
        self.textures.append(a)
        return len(self.textures)

    # ---- curves (scene.cpp:463-489) -------------------------------------------------------------
    def createCurve(self, vertex_counts, points, widths) -> int:  # noqa: N802
        points = np.ascontiguousarray(points, dtype=_F).reshape(-1, 3)
        counts = np.ascontiguousarray(vertex_counts, dtype=np.uint32).reshape(-1)
        widths = np.ascontiguousarray(widths, dtype=_F).reshape(-1)
        if len(widths):
            wstart, wcount = self._ncwidths, len(widths)
        else:
            wstart, wcount = 0xFFFFFFFF, 0xFFFFFFFF
        if self.journal is not None:
            self.journal.append(("curve", counts, points, widths))
        self.curves.append((self._nccounts, len(counts), self._ncpoints, len(points), wstart, wcount))
        self._cpoints.append(points)
        self._ccounts.append(counts)
        self._cwidths.append(widths)
        self._ncpoints += len(points)
        self._nccounts += len(counts)
        self._ncwidths += len(widths)
        return len(self.curves) - 1

    # ---- light meshes (scene.cpp:119-250) ---------------------------------------------------------
    def _create_rect_light_mesh(self) -> int:
        if self._rect_mesh < 0:
            pos = [(0.5, 0.5, 0.0), (-0.5, 0.5, 0.0), (-0.5, -0.5, 0.0), (0.5, -0.5, 0.0)]
            vb = make_vertices(pos, normals=[(0.0, 0.0, 1.0)] * 4)
            vb["uv"] = 0  # the reference leaves tangent/uv uninitialised/zero here
            self._rect_mesh = self.createMesh(vb, [0, 1, 2, 2, 3, 0])
        return self._rect_mesh

    def _create_sphere_light_mesh(self) -> int:
        if self._sphere_mesh < 0:
            seg = rings = 16
            pos, nrm, idx = [], [], []
            for i in range(rings + 1):
                # scene.cpp:160-176: float angles, ::sin / ::cos of the C library (double) narrowed to float, float products
                # (pinned against the reference's own scene.cpp by tests/test_adapter_real_headers.py)
                theta = _F(_F(i) * _F(np.pi) / _F(rings))
                st, ct = _F(np.sin(np.float64(theta))), _F(np.cos(np.float64(theta)))
                for j in range(seg + 1):
                    phi = _F(_F(_F(j) * _F(2.0)) * _F(np.pi) / _F(seg))
                    sp, cp = _F(np.sin(np.float64(phi))), _F(np.cos(np.float64(phi)))
                    p = (_F(cp * st), ct, _F(sp * st))
                    pos.append(p)
                    nrm.append(p)
            for i in range(rings):
                for j in range(seg):
                    p0 = i * (seg + 1) + j
                    p1, p2 = p0 + 1, (i + 1) * (seg + 1) + j
                    p3 = p2 + 1
                    idx += [p0, p1, p2, p2, p1, p3]
            vb = make_vertices(pos, normals=nrm)
            vb["uv"] = 0
            self._sphere_mesh = self.createMesh(vb, idx)
        return self._sphere_mesh

    def _create_disc_light_mesh(self) -> int:
        if self._disc_mesh < 0:
            pos = [(0.0, 0.0, 0.0), (1.0, 0.0, 0.0)]
            idx = []
            step = _F(2.0 * np.pi / 16)  # scene.cpp:224-231: `const float step`, `float angle` accumulated in float
            angle = _F(0.0)
            for _ in range(16):
                idx += [0, len(pos) - 1]
                angle = _F(angle + step)
                pos.append((_F(np.cos(np.float64(angle))), _F(np.sin(np.float64(angle))), 0.0))
                idx.append(len(pos) - 1)
            vb = make_vertices(pos, normals=[(0.0, 0.0, 1.0)] * len(pos))
            vb["uv"] = 0
            self._disc_mesh = self.createMesh(vb, idx)
        return self._disc_mesh

    # ---- lights (scene.cpp:306-408) ---------------------------------------------------------------
    def createLight(self, desc: UniformLightDesc) -> int:  # noqa: N802
        light_id = len(self.lights)
        if self.journal is not None:
            self.journal.append(("light", desc))
        self._in_light = True  # the light's mesh and instance are created by Scene::createLight itself
        self.lights.append(np.zeros((), dtype=LIGHT_DTYPE))
        self.light_descs.append(desc)
        self.updateLight(light_id, desc)
        if desc.type == 0:
            mesh_id = self._create_rect_light_mesh()
            s = scale_matrix(desc.width, desc.height, 1.0)
        elif desc.type == 1:
            mesh_id = self._create_disc_light_mesh()
            s = scale_matrix(desc.radius, desc.radius, desc.radius)
        elif desc.type == 2:
            mesh_id = self._create_sphere_light_mesh()
            s = scale_matrix(desc.radius, desc.radius, desc.radius)
        else:
            mesh_id = 0  # quirk Q8: the distant light instances mesh 0, whatever it is
            s = scale_matrix(desc.radius, desc.radius, desc.radius)
        transform = np.asarray(desc.xform, dtype=np.float64) @ s
        self.createInstance(SB_INSTANCE_LIGHT, mesh_id, 0xFFFFFFFF, transform, light_id)
        self._in_light = False
        return light_id

    def write_journal(self, path, width: int, height: int, spp_total: int, depth: int) -> None:
        """Serialise the API-call journal for adapter/adapter_ref_test.cpp (which replays it through the reference's own
        oka::Scene implementation)."""
        import struct

        if self.journal is None:
            raise ValueError("this scene was not built through the Scene API alone")
        cam = self.getCamera(0)

        def vec(f, arr):
            arr = np.ascontiguousarray(arr)
            f.write(struct.pack("<Q", len(arr)))
            f.write(arr.tobytes())

        with open(path, "wb") as f:
            f.write(struct.pack("<4I", width, height, spp_total, depth))
            f.write(struct.pack("<8f", *cam.position, *cam.orientation, cam.fov))
            for op in self.journal:
                if op[0] == "mesh":
                    f.write(struct.pack("<I", 1))
                    vec(f, op[1])
                    vec(f, op[2])
                elif op[0] == "instance":
                    f.write(struct.pack("<I", 2))
                    f.write(np.asarray(op[1], dtype=_F).tobytes())
                    f.write(struct.pack("<3I", op[2], op[3], op[4]))
                elif op[0] == "light":
                    d = op[1]
                    f.write(struct.pack("<Ii", 3, d.type))
                    f.write(np.asarray(d.xform, dtype=np.float64).reshape(4, 4).T.astype(_F).tobytes())
                    f.write(struct.pack("<8f", *d.color, d.intensity, d.width, d.height, d.radius, d.halfAngle))
                elif op[0] == "curve":
                    f.write(struct.pack("<I", 4))
                    vec(f, np.asarray(op[1], dtype=np.uint32))
                    vec(f, np.asarray(op[2], dtype=_F).reshape(-1))
                    vec(f, np.asarray(op[3], dtype=_F))
                elif op[0] == "material":
                    m = op[1]
                    f.write(struct.pack("<I", 5))
                    f.write(struct.pack("<I8f", int(m["model"] == _abi.SB_MATERIAL_USD_PREVIEW_SURFACE), *[float(x) for x in m["base_color"]],
                                        float(m["roughness"]), float(m["metallic"]), float(m["ior"]), float(m["clearcoat"]),
                                        float(m["clearcoat_roughness"])))
            f.write(struct.pack("<I", 0))

    def updateLight(self, light_id: int, desc: UniformLightDesc) -> None:  # noqa: N802
        l = self.lights[light_id]  # noqa: E741
        xf = np.asarray(desc.xform, dtype=np.float64)
        if desc.type == 0:
            m = (xf @ scale_matrix(desc.width, desc.height, 1.0)).astype(_F)
            corners = np.array(
                [(0.5, 0.5, 0.0, 1.0), (-0.5, 0.5, 0.0, 1.0), (-0.5, -0.5, 0.0, 1.0), (0.5, -0.5, 0.0, 1.0)], dtype=_F
            )
            l["points"] = (m @ corners.T).T
            l["type"] = 0
        elif desc.type == 1:
            m = (xf @ scale_matrix(desc.radius, desc.radius, desc.radius)).astype(_F)
            l["points"][0] = (desc.radius, 0.0, 0.0, 0.0)
            l["points"][1] = m @ np.array([0.0, 0.0, 0.0, 1.0], dtype=_F)
            l["points"][2] = m @ np.array([1.0, 0.0, 0.0, 0.0], dtype=_F)
            l["points"][3] = m @ np.array([0.0, 1.0, 0.0, 0.0], dtype=_F)
            l["normal"] = m @ np.array([0.0, 0.0, 1.0, 0.0], dtype=_F)
            l["type"] = 1
        elif desc.type == 2:
            m = xf.astype(_F)
            l["points"][0] = (desc.radius, 0.0, 0.0, 0.0)
            l["points"][1] = m @ np.array([0.0, 0.0, 0.0, 1.0], dtype=_F)
            l["type"] = 2
        else:
            m = xf.astype(_F)
            n = m @ np.array([0.0, 0.0, -1.0, 0.0], dtype=_F)
            l["normal"] = n / np.linalg.norm(n)
            l["half_angle"] = desc.halfAngle
            l["type"] = 3
        l["color"] = np.array([*desc.color, 1.0], dtype=_F) * _F(desc.intensity)

    # ---- getters ---------------------------------------------------------------------------------
    def getCamera(self, i: int = 0) -> Camera:  # noqa: N802
        return self.cameras[i]

    def arrays(self) -> dict:
        def cat(lst, dtype, shape=None):
            if not lst:
                return np.zeros((0,) + (shape or ()), dtype=dtype)
            return np.ascontiguousarray(np.concatenate(lst))

        a = {
            "vertices": cat(self._vb, VERTEX_DTYPE),
            "indices": cat(self._ib, np.uint32),
            "meshes": np.array(self.meshes, dtype=np.uint32).reshape(-1, 4).view(MESH_DTYPE).reshape(-1)
            if self.meshes else np.zeros(0, dtype=MESH_DTYPE),
            "curves": np.array(self.curves, dtype=np.uint32).reshape(-1, 6).view(CURVE_DTYPE).reshape(-1)
            if self.curves else np.zeros(0, dtype=CURVE_DTYPE),
            "curve_points": cat(self._cpoints, _F, (3,)),
            "curve_widths": cat(self._cwidths, _F),
            "curve_vertex_counts": cat(self._ccounts, np.uint32),
            "lights": np.array(self.lights, dtype=LIGHT_DTYPE) if self.lights else np.zeros(0, dtype=LIGHT_DTYPE),
            "materials": np.array(self.materials, dtype=MATERIAL_DTYPE) if self.materials else np.zeros(0, dtype=MATERIAL_DTYPE),
        }
        inst = np.zeros(len(self.instances), dtype=INSTANCE_DTYPE)
        for i, (t, ty, g, m, l) in enumerate(self.instances):  # noqa: E741
            inst[i] = (t, ty, g, m, l)
        a["instances"] = inst
        return a

    def host_bytes(self) -> int:
        """Bytes of all scene arrays a backend uploads (the H2D volume of sb_set_scene)."""
        return int(sum(v.nbytes for v in self.arrays().values()) + sum(t.nbytes for t in self.textures))

    def view(self, pinned: bool = False) -> _abi.sb_scene_view:
        """Borrowed sb_scene_view over freshly flattened arrays (kept alive on self).  pinned=True stages
        the arrays in page-locked host memory (torch) so that the upload is a true async DMA."""
        a = self.arrays()
        if pinned:
            import torch

            keep = []
            for k, arr in list(a.items()):
                if arr.nbytes == 0:
                    continue
                t = torch.empty(arr.nbytes, dtype=torch.uint8).pin_memory()
                t.numpy()[:] = np.frombuffer(arr.tobytes(), dtype=np.uint8)
                keep.append(t)
                a[k] = np.frombuffer(t.numpy().data, dtype=arr.dtype).reshape(arr.shape)
            self._keep_pinned = keep
        self._keep = a
        v = _abi.sb_scene_view()
        p = _abi.np_ptr
        v.vertices, v.num_vertices = p(a["vertices"]), len(a["vertices"])
        v.indices, v.num_indices = p(a["indices"]), len(a["indices"])
        v.meshes, v.num_meshes = p(a["meshes"]), len(a["meshes"])
        v.curves, v.num_curves = p(a["curves"]), len(a["curves"])
        v.curve_points, v.num_curve_points = p(a["curve_points"]), len(a["curve_points"])
        v.curve_widths, v.num_curve_widths = p(a["curve_widths"]), len(a["curve_widths"])
        v.curve_vertex_counts, v.num_curve_vertex_counts = p(a["curve_vertex_counts"]), len(a["curve_vertex_counts"])
        v.instances, v.num_instances = p(a["instances"]), len(a["instances"])
        v.lights, v.num_lights = p(a["lights"]), len(a["lights"])
        v.materials, v.num_materials = p(a["materials"]), len(a["materials"])
        if self.textures:
            recs = (_abi.sb_texture * len(self.textures))()
            for i, t in enumerate(self.textures):
                recs[i].pixels = t.ctypes.data
                recs[i].width, recs[i].height = t.shape[1], t.shape[0]
            self._keep_tex = recs
            import ctypes as _C

            v.textures, v.num_textures = _C.cast(recs, _C.c_void_p), len(self.textures)
        return v

    def stats(self) -> dict:
        tri = sum(self.meshes[g][1] // 3 for (_, ty, g, _, _) in self.instances if ty in (SB_INSTANCE_MESH, SB_INSTANCE_LIGHT))
        return {"instances": len(self.instances), "triangles": tri, "lights": len(self.lights), "materials": len(self.materials)}
