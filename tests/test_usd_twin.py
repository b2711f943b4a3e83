"""`.usda` twins (strelka_b200/usd_twin.py): Scene -> text -> prims -> Hydra-delegate flattening -> Scene' must
render like the original scene (CPU oracle), i.e. the writer, the reader and the restated flattening agree."""
import numpy as np

from conftest import rel_rmse
from oracle import pyoracle
from strelka_b200 import _abi, usd_twin
from strelka_b200.scenes import make_cornell, make_hair
from util import random_scene


def _round_trip(scene, settings, tmp_path, w, h):
    path = str(tmp_path / "twin.usda")
    usd_twin.write_usda(scene, path, settings, w, h)
    doc = usd_twin.read_usda(path)
    return doc, usd_twin.ingest(doc)


def test_cornell_twin_renders_bit_identically(tmp_path):
    s, st, _ = make_cornell(32, 32, 4)
    doc, s2 = _round_trip(s, st, tmp_path, 32, 32)
    assert doc["layer"] == {"width": 32, "height": 32, "spp": 1, "sppTotal": 4, "depth": 4, "rectLightSamplingMethod": 0}
    assert len(s2.instances) == len(s.instances) and len(s2.lights) == len(s.lights)
    # the delegate un-indexes meshes (Mesh.cpp:143-145) and gives every unbound mesh its own default_material
    for inst in s2.instances:
        if inst[1] == _abi.SB_INSTANCE_MESH:
            _, index_count, _, vertex_count = s2.meshes[inst[2]]
            assert index_count == vertex_count
    assert len(s2.materials) == sum(1 for i in s2.instances if i[1] == _abi.SB_INSTANCE_MESH)
    a = pyoracle.OracleScene(s).render(st, 32, 32, 4)[0]
    b = pyoracle.OracleScene(s2).render(st, 32, 32, 4)[0]
    assert np.array_equal(a, b)


def test_random_scene_twin_with_preview_materials_and_light_types(tmp_path):
    s, st = random_scene(seed=4, lights=("rect", "sphere", "disc"))
    st.setAs("render/pt/sppTotal", 4)
    _, s2 = _round_trip(s, st, tmp_path, 32, 24)
    kinds = sorted(int(m["model"]) for m in s2.materials)
    assert _abi.SB_MATERIAL_USD_PREVIEW_SURFACE in kinds and _abi.SB_MATERIAL_DIFFUSE in kinds
    assert [int(l["type"]) for l in s2.lights] == [int(l["type"]) for l in s.lights]
    a = pyoracle.OracleScene(s).render(st, 32, 24, 4)[0]
    b = pyoracle.OracleScene(s2).render(st, 32, 24, 4)[0]
    assert rel_rmse(a, b) < 1e-6


def test_distant_light_round_trip_keeps_direction_and_radiance(tmp_path):
    s, st = random_scene(seed=6, lights=("distant",))
    _, s2 = _round_trip(s, st, tmp_path, 16, 16)
    l1, l2 = s.lights[0], s2.lights[0]
    assert int(l2["type"]) == 3
    np.testing.assert_allclose(l2["normal"], l1["normal"], atol=1e-6)
    np.testing.assert_allclose(l2["color"], l1["color"], rtol=1e-5)
    np.testing.assert_allclose(float(l2["half_angle"]), float(l1["half_angle"]), rtol=1e-5)


def test_curve_twin_recreates_phantom_points(tmp_path):
    s, st, (w, h) = make_hair(24, 24, 2, n_strands=40, segments=6)
    _, s2 = _round_trip(s, st, tmp_path, w, h)
    a1, a2 = s.arrays(), s2.arrays()
    assert np.array_equal(a1["curve_vertex_counts"], a2["curve_vertex_counts"])
    np.testing.assert_allclose(a2["curve_points"], a1["curve_points"], atol=1e-6)
    np.testing.assert_allclose(a2["curve_widths"], a1["curve_widths"], rtol=1e-6)
    a = pyoracle.OracleScene(s).render(st, w, h, 2)[0]
    b = pyoracle.OracleScene(s2).render(st, w, h, 2)[0]
    assert rel_rmse(a, b) < 1e-3


def test_ingest_triangulates_quads_and_replaces_vertex_normals_q21(tmp_path):
    # a unit cube authored with quads and (wrong) vertex normals: the delegate fan-triangulates (HdMeshUtil) and
    # replaces vertex-interpolated normals by smooth normals (quirk Q21), so the bogus normals must not survive
    path = tmp_path / "cube.usda"
    path.write_text("""#usda 1.0
(
    upAxis = "Y"
)
def Xform "World"
{
    def Mesh "cube"
    {
        matrix4d xformOp:transform = ( (1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1) )
        uniform token[] xformOpOrder = ["xformOp:transform"]
        int[] faceVertexCounts = [4, 4, 4, 4, 4, 4]
        int[] faceVertexIndices = [0, 1, 3, 2, 4, 6, 7, 5, 0, 4, 5, 1, 2, 3, 7, 6, 0, 2, 6, 4, 1, 5, 7, 3]
        point3f[] points = [(-1, -1, -1), (-1, -1, 1), (-1, 1, -1), (-1, 1, 1), (1, -1, -1), (1, -1, 1), (1, 1, -1), (1, 1, 1)]
        normal3f[] normals = [(0, 1, 0), (0, 1, 0), (0, 1, 0), (0, 1, 0), (0, 1, 0), (0, 1, 0), (0, 1, 0), (0, 1, 0)] (
            interpolation = "vertex"
        )
    }
}
""")
    s2 = usd_twin.ingest(usd_twin.read_usda(str(path)))
    a = s2.arrays()
    assert len(a["vertices"]) == 36 and len(a["indices"]) == 36  # 12 triangles, un-indexed
    from strelka_b200.scene import unpack_normal

    n = unpack_normal(a["vertices"]["normal"])
    p = a["vertices"]["pos"]
    expect = p / np.linalg.norm(p, axis=1, keepdims=True)  # smooth normals of a cube point along the diagonals
    assert np.abs(n - expect).max() < 4e-3  # 10-bit quantisation
    # outward winding is preserved by the fan
    tri = p.reshape(12, 3, 3)
    face_n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    assert np.all((face_n * tri.mean(axis=1)).sum(axis=1) > 0)
