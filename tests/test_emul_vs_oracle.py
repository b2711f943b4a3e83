"""The backend's device logic, instantiated on the host (tests/emul), against the CPU oracle.

CPU only: catches logic errors in the BVH builder, the CWBVH traversal and the wavefront
shade/NEE/accumulate code before any GPU time is spent.  Both sides are compiled without FP
contraction, so everything except libm calls agrees bit for bit."""
import numpy as np
import pytest

from conftest import bits_to_f32, rel_rmse, ulp_diff
from oracle import pyoracle
from strelka_b200 import _abi
from strelka_b200.scenes import make_cornell
import pyemul
from util import random_rays, random_scene


def test_device_sampler_bit_exact(golden):
    i = np.array(golden["sampler_in"], dtype=np.uint32).reshape(-1, 6)
    out = pyemul.sampler(i[:, 0], i[:, 1], i[:, 2], i[:, 3], i[:, 4], i[:, 5])
    assert np.array_equal(out.view(np.uint32), np.array(golden["sampler_out"], dtype=np.uint32))


@pytest.mark.parametrize("key,ltype,method", [
    ("light_sample_rect_uniform", 0, 0), ("light_sample_rect_sphquad", 0, 1),
    ("light_sample_sphere", 2, 0), ("light_sample_distant", 3, 0)])
def test_device_lights_match_reference(golden, key, ltype, method):
    raw = np.array(golden["light_structs"], dtype=np.uint32).reshape(-1, 28)
    lights = np.zeros(len(raw), dtype=_abi.LIGHT_DTYPE)
    lights.view(np.uint32).reshape(-1, 28)[:] = raw
    hp = bits_to_f32(golden["light_hit_points"]).reshape(-1, 3)
    u = bits_to_f32(golden["light_u"]).reshape(-1, 2)
    sel = lights["type"] == ltype
    ref = bits_to_f32(golden[key]).reshape(-1, 12)
    got = pyemul.light_sample(lights[sel], hp[sel], u[sel], method)
    assert ulp_diff(got, ref).max() <= 1


def test_camera_matrices_match_oracle():
    s, st, _ = make_cornell(64, 48, 4)
    cam = s.getCamera(0)
    cam.look_at((0.3, -0.2, 2.0), (0.1, 0.0, 0.0))
    o = pyoracle.OracleScene(s)
    c2v_o, v2w_o = o.camera_matrices(64, 48)
    c2v_e, v2w_e = pyemul.camera(cam.view_glm(), cam.fov, 64 / 48)
    assert np.array_equal(c2v_o, c2v_e)
    np.testing.assert_allclose(v2w_o, v2w_e, rtol=0, atol=1e-7)


@pytest.mark.parametrize("scene_kind", ["cornell", "random"])
def test_traversal_hits_bit_exact(scene_kind):
    if scene_kind == "cornell":
        s, _, _ = make_cornell(32, 32, 1)
        rays = random_rays(20000, seed=3, extent=0.45)
    else:
        s, _ = random_scene(seed=5)
        rays = random_rays(30000, seed=4)
    o, e = pyoracle.OracleScene(s), pyemul.EmulScene(s)
    assert o.info()["triangles"] == e.info()["triangles"]
    ho, (he, stats) = o.trace(rays, 0), e.trace(rays, 0, with_stats=True)
    assert stats[3] == 0, "traversal stack overflow"
    for f in ("kind", "t", "u", "v", "prim", "instance"):
        assert np.array_equal(ho[f], he[f]), f
    assert (ho["kind"] == 1).mean() > 0.5
    # any-hit / mask 3: light geometry is invisible to shadow rays
    rays[:, 7] = np.random.default_rng(0).uniform(0.1, 2.0, len(rays)).astype(np.float32)
    assert np.array_equal(o.trace(rays, 1)["kind"], e.trace(rays, 1)["kind"])


def test_wide_bvh_structure_and_fill():
    # every triangle referenced exactly once, inner mask == meta bytes, and the dynamic-programming cut
    # (bvh.cuh: collapse_dp_node) fills the 8-wide nodes: a greedy cut left this scene below 5 children per node
    s, _ = random_scene(seed=3, n_meshes=8, n_instances=40)
    e = pyemul.EmulScene(s)
    info, st = e.info(), e.bvh_stats()
    assert st["errors"] == 0
    assert st["nodes"] == info["tri_nodes"] and st["prim_refs"] == info["triangles"]
    assert st["inner_children"] == st["nodes"] - 1
    assert (st["inner_children"] + st["leaf_slots"]) / st["nodes"] >= 6.0
    assert st["inner_children"] + st["leaf_slots"] + st["empty_slots"] == 8 * st["nodes"]
    s, _, _ = make_cornell(16, 16, 1)
    st = pyemul.EmulScene(s).bvh_stats()
    assert st["errors"] == 0 and st["prim_refs"] == 36 and st["nodes"] == 3


def test_empty_and_tiny_scenes():
    from strelka_b200.scene import Scene
    from strelka_b200.scenes.common import make_quad_mesh
    s = Scene()
    e = pyemul.EmulScene(s)
    assert e.info()["tri_nodes"] == 0
    assert (e.trace(random_rays(16))["kind"] == 0).all()
    # one mesh, one instance, two triangles
    vb, ib = make_quad_mesh((-1, 0, 1), (1, 0, 1), (1, 0, -1), (-1, 0, -1))
    m = s.createMesh(vb, ib)
    s.createInstance(_abi.SB_INSTANCE_MESH, m, 0, np.eye(4))
    o, e = pyoracle.OracleScene(s), pyemul.EmulScene(s)
    rays = random_rays(2000, seed=1)
    ho, he = o.trace(rays), e.trace(rays)
    for f in ("kind", "t", "u", "v", "prim", "instance"):
        assert np.array_equal(ho[f], he[f]), f
    # a single triangle
    s2 = Scene()
    m = s2.createMesh(vb[:3], ib[:3])
    s2.createInstance(_abi.SB_INSTANCE_MESH, m, 0, np.eye(4))
    ho, he = pyoracle.OracleScene(s2).trace(rays), pyemul.EmulScene(s2).trace(rays)
    assert np.array_equal(ho["t"], he["t"]) and np.array_equal(ho["kind"], he["kind"])


def test_invalid_scene_is_rejected():
    from strelka_b200.scene import Scene
    from strelka_b200.scenes.common import make_quad_mesh
    s = Scene()
    vb, ib = make_quad_mesh((-1, 0, 1), (1, 0, 1), (1, 0, -1), (-1, 0, -1))
    ib = ib.copy()
    ib[2] = 99  # index beyond the mesh's vertex count
    m = s.createMesh(vb, ib)
    s.createInstance(_abi.SB_INSTANCE_MESH, m, 0, np.eye(4))
    with pytest.raises(RuntimeError, match="index beyond"):
        pyemul.EmulScene(s)


@pytest.mark.parametrize("rect_method", [0, 1])
def test_cornell_render_matches_oracle(rect_method):
    s, st, _ = make_cornell(48, 48, 8, rect_method=rect_method)
    o, e = pyoracle.OracleScene(s), pyemul.EmulScene(s)
    img_o, _, sub, cnt_o = o.render(st, 48, 48, 8)
    img_e, _, cnt_e = e.render(st, 48, 48, 8, chunk_max=3)
    assert sub == 8
    assert cnt_o["paths"] == cnt_e["paths"] and cnt_o["radiance_rays"] == cnt_e["radiance_rays"]
    assert cnt_e["shadow_rays"] <= cnt_o["shadow_rays"]  # zero-contribution shadow rays are not traced
    assert rel_rmse(img_e, img_o) < 1e-5
    assert img_o[..., :3].mean() > 0.01


def test_random_scene_render_matches_oracle():
    s, st = random_scene(seed=7)
    st.setAs("render/pt/sppTotal", 4)
    st.setAs("render/pt/depth", 6)  # russian roulette active for bounces 4,5 (quirk Q9)
    o, e = pyoracle.OracleScene(s), pyemul.EmulScene(s)
    img_o, _, _, cnt_o = o.render(st, 40, 30, 4)
    img_e, _, cnt_e = e.render(st, 40, 30, 4, chunk_max=2)
    assert cnt_o["radiance_rays"] == cnt_e["radiance_rays"]
    assert rel_rmse(img_e, img_o) < 1e-5


def test_hair_material_render_matches_oracle():
    """C4 with the hair fibre BSDF (float device code in the host emulation vs the oracle's double-precision restatement
    from the papers): transmission events toggle `inside`, russian roulette is live at depth 6"""
    from strelka_b200.scenes import make_hair

    s, st, _ = make_hair(48, 48, 4, depth=6, n_strands=1500, segments=6, material="hair")
    o, e = pyoracle.OracleScene(s), pyemul.EmulScene(s)
    img_o, _, _, cnt_o = o.render(st, 48, 48, 4)
    img_e, _, cnt_e = e.render(st, 48, 48, 4, chunk_max=2)
    assert img_o[..., :3].mean() > 1e-3
    assert abs(cnt_o["radiance_rays"] / cnt_e["radiance_rays"] - 1.0) < 1e-3
    assert rel_rmse(img_e, img_o) < 1e-3
    assert abs(img_e[..., :3].mean() / img_o[..., :3].mean() - 1.0) < 5e-3


def test_debug_normals_exact():
    s, st = random_scene(seed=2)
    st.setAs("render/pt/debug", 1)
    o, e = pyoracle.OracleScene(s), pyemul.EmulScene(s)
    img_o, _, _, _ = o.render(st, 40, 30, 1)
    img_e, _, _ = e.render(st, 40, 30, 1)
    assert np.array_equal(img_o[..., :3], img_e[..., :3])


@pytest.mark.parametrize("debug", [2, 3])
def test_aov_views_match_oracle(debug):
    # diffuse / specular AOVs (OptixRender.cu:157-221): per-AOV sample counters, image = that AOV
    s, st = random_scene(seed=5)
    st.setAs("render/pt/sppTotal", 6)
    st.setAs("render/pt/debug", debug)
    o, e = pyoracle.OracleScene(s), pyemul.EmulScene(s)
    aov = np.zeros((30, 40, 10), dtype=np.float32)
    img_o, _, sub, _ = o.render(st, 40, 30, 6, aov=aov)
    img_e, _, _ = e.render(st, 40, 30, 6, chunk_max=4)
    assert sub == 6
    assert img_o[..., :3].max() > 0.0
    counts = aov[..., 8 if debug == 2 else 9]
    assert counts.max() > 0 and counts.min() == 0  # some pixels never saw this event: they stay black
    assert np.all(img_e[counts == 0][..., :3] == 0.0)
    assert rel_rmse(img_e, img_o) < 1e-5


@pytest.mark.parametrize("debug", [0, 1, 3])
def test_fused_path_form_equals_wavefront_form(debug):
    # k_path_fused (small scenes) runs whole paths with the state in registers: bit-identical to the queue form
    s, st = random_scene(seed=9)
    st.setAs("render/pt/sppTotal", 5)
    st.setAs("render/pt/depth", 6)
    st.setAs("render/pt/debug", debug)
    e = pyemul.EmulScene(s)
    n = 1 if debug == 1 else 5
    a, Sa, ca = e.render(st, 37, 26, n, chunk_max=2)
    b, Sb, cb = e.render(st, 37, 26, n, chunk_max=2, fused=True)
    assert np.array_equal(a, b) and np.array_equal(Sa, Sb)
    if debug != 1:
        assert ca == cb


def test_progressive_equals_batched():
    # 6 launches one by one == one call with 6 samples in chunks of 4 (sum form is order independent)
    s, st, _ = make_cornell(32, 32, 6)
    e = pyemul.EmulScene(s)
    a, _, _ = e.render(st, 32, 32, 6, chunk_max=4)
    S = None
    for k in range(6):
        b, S, _ = e.render(st, 32, 32, 1, subframe=k, chunk_max=1, S=S)
    assert np.array_equal(a, b)


def test_sample_stride_shards_sum_to_whole():
    # multi-GPU sharding property: S over strided sample subsets adds up to the full S
    s, st, _ = make_cornell(32, 32, 8)
    e = pyemul.EmulScene(s)
    _, S_full, _ = e.render(st, 32, 32, 8, chunk_max=8)
    S_sum = np.zeros_like(S_full)
    for g in range(2):
        st.setAs("render/b200/sampleOffset", g)
        st.setAs("render/b200/sampleStride", 2)
        _, S_g, _ = e.render(st, 32, 32, 4, chunk_max=4)
        S_sum += S_g
    np.testing.assert_allclose(S_sum, S_full, rtol=2e-6, atol=1e-9)


def test_octant_permutation_table_of_the_fixed_bit_hit_mask():
    """traverse.cuh: inner-child hit bits are gathered by SLOT and moved to traversal order (bit s -> bit s ^ octinv) by a
    256-entry table per octant.  The device reads the table, the host emulation computes the permutation: both must
    agree, and both must be the XOR permutation (an involution that keeps the population count)."""
    from emul import pyemul

    bad, t = pyemul.octant_perm_table()
    assert bad == 0
    x = np.arange(256)
    for o in range(8):
        want = np.zeros(256, np.int64)
        for b in range(8):
            want |= ((x >> b) & 1) << (b ^ o)
        assert np.array_equal(t[o], want)
        assert np.array_equal(t[o][t[o]], x)  # applying it twice is the identity
