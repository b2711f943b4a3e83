"""`.usda` twins of the synthetic scenes, and a restatement of the Hydra delegate's flattening.

Why: the reference reads scenes only through OpenUSD (src/HdStrelka), which is not available here.  A twin is the
same scene written as plain-text USD, so that a maintainer with a Strelka build can render exactly what this
repository benchmarks (`write_usda`) -- the one way left to pin the integrator-level parity that is unpinned today --
and so that the flattening contract of the delegate is executable here (`ingest`):

    Scene --write_usda--> text --read_usda--> prims --ingest--> Scene'      (tests: Scene' renders like Scene)

`ingest` follows, line by line where it matters for the arrays a backend sees:
  * Mesh.cpp:123-180       un-indexed triangles (3 vertices per face), tangent from the normal (computeTangent)
  * Mesh.cpp:251-279       face-varying normals are used as authored (quirk Q21: vertex normals would be replaced)
  * RenderPass.cpp:69-126  packNormal / packUV (v flipped), one oka mesh + one instance per (prim, instance) (quirk Q16)
  * RenderPass.cpp:218-242 a mesh without material binding gets its own default_material with diffuse_color =
                           constant displayColor (white if absent)
  * BasisCurves.cpp:189-231 phantom end points once per prim, widths / 2 (quirk Q14)
  * Light.cpp:110-208      intensity * 2^exposure; disc/sphere radius * xform[0][0]; distant: halfAngle = angle/2,
                           intensity / (pi sin^2 halfAngle)
  * Camera.cpp:65-106      position = translation of the camera transform, orientation = conjugate of its rotation,
                           fov = vertical field of view in degrees
Prims are visited in path order, as HdRenderIndex::GetRprimIds / GetSprimSubtree return them; the writer names prims
so that this order is the original instance order.  Of OpenUSD itself only the two published mesh utilities the delegate calls are restated (`triangulate` = HdMeshUtil's
fan triangulation, `smooth_normals` = Hd_SmoothNormals); UsdImaging, instancers and the MaterialX translation are not:
the twins only contain triangles with face-varying normals, direct transforms and UsdPreviewSurface constants.  Parity of this module against the real
delegate is therefore unpinned; it documents the contract and keeps the twins honest.
"""
from __future__ import annotations

import math
import re

import numpy as np

from . import _abi
from .camera import Camera
from .scene import Scene, UniformLightDesc, make_vertices

_F = np.float32


# ---------------------------------------------------------------------------------------------- writer
def _f(x) -> str:
    return repr(float(np.float32(x))) if np.isfinite(x) else "0"


def _tuple(v) -> str:
    return "(" + ", ".join(_f(x) for x in v) + ")"


def _array(vs) -> str:
    return "[" + ", ".join(_tuple(v) for v in vs) + "]"


def _matrix(m) -> str:
    """math-convention 4x4 (p' = M p) -> USD row-vector matrix4d (translation in the last row)."""
    t = np.asarray(m, dtype=np.float64).T
    return "( " + ", ".join("(" + ", ".join(repr(float(x)) for x in row) + ")" for row in t) + " )"


def _requantisable_normals(packed) -> np.ndarray:
    """Normals that packNormal maps back onto the same 10-bit codes (cell centres, not cell edges)."""
    v = np.asarray(packed, dtype=np.uint32)
    q = np.stack([v & 0x3FF, (v >> 10) & 0x3FF, (v >> 20) & 0xFFF], axis=-1).astype(np.float64)
    return ((q + 0.5) / 511.99999 * 2.0 - 1.0).astype(_F)


def write_usda(scene: Scene, path: str, settings=None, width: int = 0, height: int = 0) -> None:
    a = scene.arrays()
    out = ["#usda 1.0", "(", '    upAxis = "Y"', "    metersPerUnit = 1", '    doc = "strelka_b200 synthetic scene twin"']
    if settings is not None:
        st = settings.to_sb_settings()
        out.append(f'    customLayerData = {{ int width = {width}; int height = {height}; int spp = {st.spp}; int sppTotal = {st.spp_total}; '
                   f"int depth = {st.depth}; int rectLightSamplingMethod = {st.rect_light_sampling_method} }}")
    out += [")", "", 'def Xform "World"', "{"]
    mats = a["materials"]
    used_preview = set()
    n = 0
    for inst in scene.instances:
        xf, type_, geom, mat, light = inst
        m = np.asarray(xf, dtype=np.float64).reshape(4, 4).T  # back to math convention
        name = f"i{n:07d}"
        n += 1
        if type_ == _abi.SB_INSTANCE_MESH:
            first, count, vbo, nv = scene.meshes[geom]
            idx = a["indices"][first:first + count].astype(np.int64) + vbo
            vb = a["vertices"][idx]
            ntri = count // 3
            out += [f'    def Mesh "{name}"', "    {", f"        matrix4d xformOp:transform = {_matrix(m)}",
                    '        uniform token[] xformOpOrder = ["xformOp:transform"]',
                    "        int[] faceVertexCounts = [" + ", ".join(["3"] * ntri) + "]",
                    "        int[] faceVertexIndices = [" + ", ".join(str(i) for i in range(3 * ntri)) + "]",
                    f"        point3f[] points = {_array(vb['pos'])}",
                    f"        normal3f[] normals = {_array(_requantisable_normals(vb['normal']))} (", '            interpolation = "faceVarying"', "        )"]
            md = mats[mat] if mat < len(mats) else None
            if md is not None and int(md["model"]) == _abi.SB_MATERIAL_USD_PREVIEW_SURFACE:
                used_preview.add(mat)
                out.append(f"        rel material:binding = </World/Materials/m{mat:05d}>")
            else:
                color = md["base_color"] if md is not None else (1.0, 1.0, 1.0)
                out += [f"        color3f[] primvars:displayColor = [{_tuple(color)}] (", '            interpolation = "constant"', "        )"]
            out.append("    }")
        elif type_ == _abi.SB_INSTANCE_CURVE:
            cs, cn, ps, pn, ws, wn = scene.curves[geom]
            pts = a["curve_points"][ps:ps + pn]
            counts = a["curve_vertex_counts"][cs:cs + cn]
            out += [f'    def BasisCurves "{name}"', "    {", f"        matrix4d xformOp:transform = {_matrix(m)}",
                    '        uniform token[] xformOpOrder = ["xformOp:transform"]', '        uniform token type = "cubic"',
                    '        uniform token basis = "bspline"', "        int[] curveVertexCounts = [" + ", ".join(str(int(c)) for c in counts) + "]",
                    f"        point3f[] points = {_array(pts[1:-1])}"]  # the delegate re-creates the two phantom points (Q14)
            if wn not in (0, 0xFFFFFFFF):
                w = a["curve_widths"][ws:ws + wn][1:-1] * _F(2.0)
                out.append("        float[] widths = [" + ", ".join(_f(x) for x in w) + "]")
            if mat < len(mats):
                used_preview.add(mat)
                out.append(f"        rel material:binding = </World/Materials/m{mat:05d}>")
            out.append("    }")
        elif type_ == _abi.SB_INSTANCE_LIGHT:
            d = scene.light_descs[light]
            kind = {0: "RectLight", 1: "DiskLight", 2: "SphereLight", 3: "DistantLight"}[d.type]
            out += [f'    def {kind} "{name}"', "    {", f"        matrix4d xformOp:transform = {_matrix(d.xform)}",
                    '        uniform token[] xformOpOrder = ["xformOp:transform"]', f"        color3f inputs:color = {_tuple(d.color)}",
                    "        float inputs:exposure = 0"]
            if d.type == 0:
                out += [f"        float inputs:intensity = {_f(d.intensity)}", f"        float inputs:width = {_f(d.width)}", f"        float inputs:height = {_f(d.height)}"]
            elif d.type in (1, 2):
                sx = float(np.asarray(d.xform, dtype=np.float64)[0, 0])
                out += [f"        float inputs:intensity = {_f(d.intensity)}", f"        float inputs:radius = {_f(d.radius / sx if sx else d.radius)}"]
            else:
                s2 = math.pi * math.sin(d.halfAngle) ** 2
                out += [f"        float inputs:intensity = {_f(d.intensity * s2)}", f"        float inputs:angle = {_f(math.degrees(2.0 * d.halfAngle))}"]
            out.append("    }")
    cam = scene.getCamera(0)
    cam.updateViewMatrix()
    c2w = np.linalg.inv(cam.view)
    aperture = 24.0
    focal = aperture / (2.0 * math.tan(math.radians(cam.fov) * 0.5))
    out += ['    def Camera "camera"', "    {", f"        matrix4d xformOp:transform = {_matrix(c2w)}", '        uniform token[] xformOpOrder = ["xformOp:transform"]',
            f"        float focalLength = {_f(focal)}", f"        float verticalAperture = {_f(aperture)}",
            f"        float horizontalAperture = {_f(aperture * (width / height if width and height else 1.0))}",
            f"        float2 clippingRange = ({_f(cam.znear)}, {_f(cam.zfar)})", "    }"]
    out += ['    def Scope "Materials"', "    {"]
    for mi in sorted(used_preview):
        md = mats[mi]
        out += [f'        def Material "m{mi:05d}"', "        {", f"            token outputs:surface.connect = </World/Materials/m{mi:05d}/Shader.outputs:surface>",
                '            def Shader "Shader"', "            {", '                uniform token info:id = "UsdPreviewSurface"',
                f"                int inputs:strelka_model = {int(md['model'])}",
                f"                color3f inputs:diffuseColor = {_tuple(md['base_color'])}", f"                float inputs:roughness = {_f(md['roughness'])}",
                f"                float inputs:metallic = {_f(md['metallic'])}", f"                float inputs:ior = {_f(md['ior'])}",
                f"                float inputs:opacity = {_f(md['opacity'])}", f"                float inputs:clearcoat = {_f(md['clearcoat'])}",
                f"                float inputs:clearcoatRoughness = {_f(md['clearcoat_roughness'])}",
                f"                color3f inputs:specularColor = {_tuple(md['specular_color'])}",
                f"                int inputs:useSpecularWorkflow = {int(md['use_specular_workflow'])}", "                token outputs:surface", "            }", "        }"]
    out += ["    }", "}", ""]
    with open(path, "w") as f:
        f.write("\n".join(out))


# ---------------------------------------------------------------------------------------------- reader
_TOKEN = re.compile(r'\s*(?:(#[^\n]*)|("(?:[^"\\]|\\.)*")|(<[^>]*>)|([{}()\[\]=,;])|([^\s{}()\[\]=,;"<>]+))')


def _tokens(text):
    pos = 0
    while True:
        m = _TOKEN.match(text, pos)
        if not m:
            return
        pos = m.end()
        if m.group(1) is None:
            yield m.group(0).strip()


def _parse_value(tok, i):
    t = tok[i]
    if t in "([":
        close = ")" if t == "(" else "]"
        vals = []
        i += 1
        while tok[i] != close:
            if tok[i] == ",":
                i += 1
                continue
            v, i = _parse_value(tok, i)
            vals.append(v)
        return vals, i + 1
    if t.startswith('"'):
        return t[1:-1], i + 1
    if t.startswith("<"):
        return t, i + 1
    try:
        return float(t), i + 1
    except ValueError:
        return t, i + 1


def _parse_block(tok, i):
    """tok[i] is just after '{'.  Returns (attrs, children, index after the closing '}')."""
    attrs, children = {}, []
    while tok[i] != "}":
        if tok[i] == "def":
            typ, name = tok[i + 1], tok[i + 2][1:-1]
            i += 3
            if tok[i] == "(":  # prim metadata
                _, i = _parse_value(tok, i)
            assert tok[i] == "{"
            a, c, i = _parse_block(tok, i + 1)
            children.append({"type": typ, "name": name, "attrs": a, "children": c})
            continue
        # attribute: [uniform|custom] type name = value [( metadata )]
        j = i
        while tok[j] not in ("=", "}", "def"):
            j += 1
        if tok[j] != "=":  # declaration without a value (`token outputs:surface`)
            i = j
            continue
        name = tok[j - 1]
        val, k = _parse_value(tok, j + 1)
        meta = {}
        if k < len(tok) and tok[k] == "(":
            k += 1
            while tok[k] != ")":
                if tok[k + 1] == "=":
                    mv, k2 = _parse_value(tok, k + 2)
                    meta[tok[k]] = mv
                    k = k2
                else:
                    k += 1
            k += 1
        attrs[name] = (val, meta)
        i = k
    return attrs, children, i + 1


def read_usda(path: str) -> dict:
    """Tiny reader for the subset `write_usda` emits (not a USD parser)."""
    text = open(path).read()
    assert text.startswith("#usda 1.0")
    tok = list(_tokens(text[len("#usda 1.0"):]))
    i = 0
    layer = {}
    if tok[0] == "(":
        depth, i = 1, 1
        while depth:
            if tok[i] == "(":
                depth += 1
            elif tok[i] == ")":
                depth -= 1
            i += 1
        hdr = " ".join(tok[: i])
        for k in ("width", "height", "spp", "sppTotal", "depth", "rectLightSamplingMethod"):
            m = re.search(rf"int {k} = (\d+)", hdr)
            if m:
                layer[k] = int(m.group(1))
    tok = tok[i:] + ["}"]
    _, children, _ = _parse_block(tok, 0)
    return {"layer": layer, "prims": children}


# ---------------------------------------------------------------------------------------------- ingest
def _usd_matrix(val) -> np.ndarray:
    """USD row-vector matrix -> math convention (this is the `xform[i][j] = transform[i][j]` copy into glm's
    column-major storage of RenderPass.cpp:117-124: a transpose in disguise)."""
    return np.asarray(val, dtype=np.float64).reshape(4, 4).T


def _compute_tangent(n: np.ndarray) -> np.ndarray:
    """computeTangent, Mesh.cpp:146-160."""
    n = n.astype(_F)
    c1 = np.cross(n, np.array([1.0, 0.0, 0.0], dtype=_F)).astype(_F)
    c2 = np.cross(n, np.array([0.0, 1.0, 0.0], dtype=_F)).astype(_F)
    l1 = (c1 * c1).sum(axis=-1, keepdims=True)
    l2 = (c2 * c2).sum(axis=-1, keepdims=True)
    t = np.where(l1 > l2, c1, c2)
    ln = np.sqrt((t * t).sum(axis=-1, keepdims=True))
    return (t / np.maximum(ln, _F(1e-30))).astype(_F)


def triangulate(counts, indices, left_handed: bool = False):
    """HdMeshUtil::ComputeTriangleIndices / ComputeTriangulatedFaceVaryingPrimvar (OpenUSD pxr/imaging/hd/meshUtil.cpp,
    not in the reference tree): fan (0, i+1, i+2) per face, winding flipped for leftHanded orientation, faces with
    fewer than 3 vertices dropped.  Returns (vertex indices [T,3], face-varying corner indices [T,3])."""
    counts = np.asarray(counts, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    tri_v, tri_c = [], []
    base = 0
    for n in counts:
        for i in range(max(int(n) - 2, 0)):
            c = (base, base + i + 2, base + i + 1) if left_handed else (base, base + i + 1, base + i + 2)
            tri_c.append(c)
            tri_v.append(tuple(indices[list(c)]))
        base += int(n)
    return np.asarray(tri_v, dtype=np.int64).reshape(-1, 3), np.asarray(tri_c, dtype=np.int64).reshape(-1, 3)


def smooth_normals(points, counts, indices, left_handed: bool = False) -> np.ndarray:
    """Hd_SmoothNormals::ComputeSmoothNormals (pxr/imaging/hd/smoothNormals.cpp): per vertex, the normalised sum over
    its incident face corners of cross(next - curr, prev - curr) -- area- and angle-weighted, float32 accumulation."""
    points = np.asarray(points, dtype=_F).reshape(-1, 3)
    counts = np.asarray(counts, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    n = np.zeros_like(points)
    base = 0
    for cnt in counts:
        cnt = int(cnt)
        for k in range(cnt):
            cur, prv, nxt = indices[base + k], indices[base + (k - 1) % cnt], indices[base + (k + 1) % cnt]
            if left_handed:
                prv, nxt = nxt, prv
            n[cur] += np.cross(points[nxt] - points[cur], points[prv] - points[cur]).astype(_F)
        base += cnt
    ln = np.sqrt((n * n).sum(axis=1, keepdims=True))
    return (n / np.maximum(ln, _F(1e-30))).astype(_F)


def _walk(prims, prefix=""):
    for p in sorted(prims, key=lambda q: q["name"]):  # path order, as the render index returns ids
        path = prefix + "/" + p["name"]
        yield path, p
        yield from _walk(p["children"], path)


def ingest(doc: dict) -> Scene:
    """The flattening the Hydra delegate performs on the prims of `read_usda` -> oka::Scene mirror."""
    s = Scene()
    prims = dict(_walk(doc["prims"]))
    material_index: dict[str, int] = {}

    def get_or_create_material(path: str) -> int:
        if path in material_index:
            return material_index[path]
        shader = next(c for c in prims[path]["children"] if c["type"] == "Shader")["attrs"]

        def val(name, default):
            return shader[name][0] if name in shader else default

        # resolveMaterial of adapter/B200Render.cpp: UsdPreviewSurface input names -> sb_material
        idx = s.addMaterial(model=int(val("inputs:strelka_model", _abi.SB_MATERIAL_USD_PREVIEW_SURFACE)), base_color=tuple(val("inputs:diffuseColor", (0.18,) * 3)),
                            roughness=val("inputs:roughness", 0.5), metallic=val("inputs:metallic", 0.0), ior=val("inputs:ior", 1.5),
                            opacity=val("inputs:opacity", 1.0), clearcoat=val("inputs:clearcoat", 0.0), clearcoat_roughness=val("inputs:clearcoatRoughness", 0.01),
                            specular_color=tuple(val("inputs:specularColor", (0.0,) * 3)), use_specular_workflow=int(val("inputs:useSpecularWorkflow", 0)))
        material_index[path] = idx
        return idx

    lights = []
    for path, p in prims.items():
        a = p["attrs"]
        if p["type"] == "Mesh":
            counts = np.asarray(a["faceVertexCounts"][0], dtype=np.int64)
            fvi = np.asarray(a["faceVertexIndices"][0], dtype=np.int64)
            points = np.asarray(a["points"][0], dtype=_F).reshape(-1, 3)
            left = "orientation" in a and a["orientation"][0] == "leftHanded"
            tri_v, tri_c = triangulate(counts, fvi, left)
            if "normals" in a and a["normals"][1].get("interpolation") == "faceVarying":
                fv = np.asarray(a["normals"][0], dtype=_F).reshape(-1, 3)
                normals = fv[tri_c.reshape(-1)]  # triangulated face-varying primvar, used as authored
            else:
                # authored vertex normals are read and then REPLACED by smooth normals (quirk Q21, Mesh.cpp:251-279)
                normals = smooth_normals(points, counts, fvi, left)[tri_v.reshape(-1)]
            uvs = None
            for key in ("primvars:st", "st"):
                if key in a:
                    st_vals = np.asarray(a[key][0], dtype=_F).reshape(-1, 2)
                    interp = a[key][1].get("interpolation", "vertex")
                    uv = st_vals[tri_c.reshape(-1)] if interp == "faceVarying" else st_vals[tri_v.reshape(-1)]
                    uvs = np.stack([uv[:, 0], _F(1.0) - uv[:, 1]], axis=1)  # v flipped, RenderPass.cpp:104
                    break
            pos = points[tri_v.reshape(-1)]  # Mesh.cpp:143-145: three new vertices per triangle
            vb = make_vertices(pos, normals=normals, tangents=_compute_tangent(normals), uvs=uvs)
            if "material:binding" in a:
                mat = get_or_create_material(a["material:binding"][0].strip("<>"))
            else:
                color = (1.0, 1.0, 1.0)
                if "primvars:displayColor" in a and a["primvars:displayColor"][1].get("interpolation") == "constant":
                    color = tuple(a["primvars:displayColor"][0][0])
                mat = s.addMaterial(model=_abi.SB_MATERIAL_DIFFUSE, base_color=color)  # a new default_material per mesh
            mesh = s.createMesh(vb, np.arange(len(pos), dtype=np.uint32))
            s.createInstance(_abi.SB_INSTANCE_MESH, mesh, mat, _usd_matrix(a["xformOp:transform"][0]))
        elif p["type"] == "BasisCurves":
            pts = np.asarray(a["points"][0], dtype=_F).reshape(-1, 3)
            p0 = pts[0] + (pts[0] - pts[1])
            pn = pts[-1] + (pts[-1] - pts[-2])
            allp = np.concatenate([p0[None], pts, pn[None]]).astype(_F)
            if "widths" in a:
                w = np.asarray(a["widths"][0], dtype=_F).reshape(-1)
                if len(w) == 1:
                    w = np.full(len(pts), w[0], dtype=_F)
                r = np.concatenate([[w[0] * _F(0.5)], w * _F(0.5)])
                r = np.concatenate([r, [r[-1]]]).astype(_F)
            else:
                r = np.zeros(0, dtype=_F)
            mat = get_or_create_material(a["material:binding"][0].strip("<>")) if "material:binding" in a else 0
            curve = s.createCurve(np.asarray(a["curveVertexCounts"][0], dtype=np.uint32), allp, r)
            s.createInstance(_abi.SB_INSTANCE_CURVE, curve, mat, _usd_matrix(a["xformOp:transform"][0]))
        elif p["type"] in ("RectLight", "DiskLight", "SphereLight", "DistantLight"):
            lights.append((path, p))
        elif p["type"] == "Camera":
            c2w = _usd_matrix(a["xformOp:transform"][0])
            cam = Camera()
            cam.position = c2w[:3, 3].copy()
            from .camera import _mat_to_quat  # noqa: PLC0415

            cam.orientation = _mat_to_quat(c2w[:3, :3].T)  # conjugate of the camera-to-world rotation
            cam.fov = math.degrees(2.0 * math.atan(a["verticalAperture"][0] / (2.0 * a["focalLength"][0])))
            cam.znear, cam.zfar = a["clippingRange"][0]
            cam.updateViewMatrix()
            s.cameras[0] = cam
    for _, p in lights:  # sprims are synced after the rprims were baked into the scene
        a = p["attrs"]
        d = UniformLightDesc()
        d.xform = _usd_matrix(a["xformOp:transform"][0])
        d.color = tuple(a["inputs:color"][0])
        d.intensity = float(_F(a["inputs:intensity"][0]) * _F(2.0) ** _F(min(max(a["inputs:exposure"][0], -50.0), 50.0)))
        if p["type"] == "RectLight":
            d.type, d.width, d.height = 0, a["inputs:width"][0], a["inputs:height"][0]
        elif p["type"] in ("DiskLight", "SphereLight"):
            d.type = 1 if p["type"] == "DiskLight" else 2
            d.radius = a["inputs:radius"][0] * float(_F(d.xform[0, 0]))
        else:
            d.type = 3
            d.halfAngle = float(a["inputs:angle"][0]) * 0.5 * (math.pi / 180.0)
            d.intensity = d.intensity / (math.pi * math.sin(d.halfAngle) ** 2)
        s.createLight(d)
    return s
