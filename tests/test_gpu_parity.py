"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the
golden vectors taken from the reference's own headers.

Tolerances (stated per test): integer work (sampler) bit-exact; traversal hits bit-exact (same
expression trees, -fmad=false vs -ffp-contract=off); images within the north-star gate of
per-pixel relative RMSE <= 1e-3 (observed ~1e-6: only libm sin/cos/acos differ)."""
import numpy as np
import pytest

from conftest import bits_to_f32, rel_rmse
from oracle import pyoracle
from strelka_b200 import _abi, Buffer, BufferDesc, BufferFormat, SharedContext
from strelka_b200.scenes import make_cornell
from util import random_rays, random_scene

pytestmark = pytest.mark.gpu


def _render(gpu_render, scene, settings, w, h, iterations, batched=True):
    r = gpu_render
    r.setScene(scene)
    r.setSharedContext(SharedContext(mSettingsManager=settings))
    r._last_settings = None
    r.reset_accumulation()
    buf = r.createBuffer(BufferDesc(w, h, BufferFormat.FLOAT4))
    if batched:
        r.render_iterations(buf, iterations)
    else:
        for _ in range(iterations):
            r.render(buf)
    img = buf.map().copy()
    buf.unmap()
    buf.destroy()
    return img


def test_sampler_bit_exact_on_device(gpu_render, golden):
    i = np.array(golden["sampler_in"], dtype=np.uint32).reshape(-1, 6)
    out = gpu_render.test_sampler(i[:, 0], i[:, 1], i[:, 2], i[:, 3], i[:, 4], i[:, 5])
    assert np.array_equal(out.view(np.uint32), np.array(golden["sampler_out"], dtype=np.uint32))


@pytest.mark.parametrize("key,ltype,method,tol", [
    ("light_sample_rect_uniform", 0, 0, 2e-6), ("light_sample_rect_sphquad", 0, 1, 2e-3),
    ("light_sample_sphere", 2, 0, 2e-6), ("light_sample_distant", 3, 0, 2e-6)])
def test_light_sampling_on_device(gpu_render, golden, key, ltype, method, tol):
    raw = np.array(golden["light_structs"], dtype=np.uint32).reshape(-1, 28)
    lights = np.zeros(len(raw), dtype=_abi.LIGHT_DTYPE)
    lights.view(np.uint32).reshape(-1, 28)[:] = raw
    hp = bits_to_f32(golden["light_hit_points"]).reshape(-1, 3)
    u = bits_to_f32(golden["light_u"]).reshape(-1, 2)
    sel = lights["type"] == ltype
    ref = bits_to_f32(golden[key]).reshape(-1, 12)
    got = gpu_render.test_light_sample(lights[sel], hp[sel], u[sel], method)
    # CUDA libm vs glibc differ by <= 2 ulp in sin/cos/acos; the spherical-rectangle sampler amplifies
    # that near grazing configurations, hence the looser bound there (relative to the value scale)
    scale = np.maximum(np.abs(ref), 1.0)
    err = np.abs(got - ref) / scale
    assert np.max(err) < tol
    assert np.percentile(err, 99) < 2e-5  # the loose bound above is for the ill-conditioned tail only


@pytest.mark.parametrize("scene_kind", ["cornell", "random"])
def test_traversal_hits_bit_exact_on_device(gpu_render, scene_kind):
    if scene_kind == "cornell":
        s, _, _ = make_cornell(32, 32, 1)
        rays = random_rays(200000, seed=3, extent=0.45)
    else:
        s, _ = random_scene(seed=5)
        rays = random_rays(200000, seed=4)
    gpu_render.setScene(s)
    hg = gpu_render.test_trace(rays, 0)
    ho = pyoracle.OracleScene(s).trace(rays, 0)
    for f in ("kind", "t", "u", "v", "prim", "instance"):
        assert np.array_equal(ho[f], hg[f]), f
    rays[:, 7] = np.random.default_rng(0).uniform(0.1, 2.0, len(rays)).astype(np.float32)
    assert np.array_equal(pyoracle.OracleScene(s).trace(rays, 1)["kind"], gpu_render.test_trace(rays, 1)["kind"])


@pytest.mark.parametrize("rect_method", [0, 1])
def test_cornell_image_matches_oracle(gpu_render, rect_method):
    w = h = 96
    s, st, _ = make_cornell(w, h, 16, rect_method=rect_method)
    img_g = _render(gpu_render, s, st, w, h, 16)
    img_o, _, _, _ = pyoracle.OracleScene(s).render(st, w, h, 16)
    assert rel_rmse(img_g, img_o) <= 1e-3  # north-star gate; expected ~1e-6
    assert abs(img_g[..., :3].mean() / img_o[..., :3].mean() - 1.0) < 5e-3


def test_random_scene_image_matches_oracle(gpu_render):
    s, st = random_scene(seed=7)
    st.setAs("render/pt/sppTotal", 8)
    st.setAs("render/pt/depth", 6)
    img_g = _render(gpu_render, s, st, 80, 60, 8)
    img_o, _, _, _ = pyoracle.OracleScene(s).render(st, 80, 60, 8)
    assert rel_rmse(img_g, img_o) <= 1e-3


def test_debug_normals_exact_on_device(gpu_render):
    s, st = random_scene(seed=2)
    st.setAs("render/pt/debug", 1)
    img_g = _render(gpu_render, s, st, 80, 60, 1, batched=False)
    img_o, _, _, _ = pyoracle.OracleScene(s).render(st, 80, 60, 1)
    # camera + traversal + attribute fetch are libm-free -> exact (analytic pin ii)
    assert np.array_equal(img_g[..., :3], img_o[..., :3])


def test_render_calls_equal_batched_iterations(gpu_render):
    s, st, _ = make_cornell(64, 64, 12)
    a = _render(gpu_render, s, st, 64, 64, 12, batched=True)
    b = _render(gpu_render, s, st, 64, 64, 12, batched=False)
    assert np.array_equal(a, b)


def test_stops_at_spp_total_like_reference(gpu_render):
    s, st, _ = make_cornell(32, 32, 4)
    a = _render(gpu_render, s, st, 32, 32, 4)
    b = _render(gpu_render, s, st, 32, 32, 9)  # 5 extra calls render nothing (OptixRender.cpp:989-1030)
    assert np.array_equal(a, b)
    assert gpu_render.getSharedContext().mSubframeIndex == 4


def test_counters_match_oracle(gpu_render):
    s, st, _ = make_cornell(64, 64, 4)
    gpu_render.reset_counters()
    _render(gpu_render, s, st, 64, 64, 4)
    c = gpu_render.counters()
    _, _, _, co = pyoracle.OracleScene(s).render(st, 64, 64, 4)
    assert c["paths"] == co["paths"] and c["radiance_rays"] == co["radiance_rays"]
    assert 0 < c["shadow_rays"] <= co["shadow_rays"]
    assert c["num_triangles"] == 36 and c["bvh_nodes_tri"] >= 1


def test_sample_stride_sharding_sums_to_single_gpu_image(gpu_render):
    # the multi-GPU decomposition on one device: two strided halves, S added, resolved with n = total
    import torch

    w = h = 64
    s, st, _ = make_cornell(w, h, 8)
    full = _render(gpu_render, s, st, w, h, 8)
    total = None
    for g in range(2):
        st.setAs("render/b200/sampleOffset", g)
        st.setAs("render/b200/sampleStride", 2)
        _render(gpu_render, s, st, w, h, 8)
        assert gpu_render.getSharedContext().mSubframeIndex == 4  # each shard owns 4 of the 8 samples
        t = gpu_render.accum_tensor().clone()
        total = t if total is None else total + t
    # write the reduced sum back (what an NCCL all-reduce does in place) and resolve with n = 8
    gpu_render.accum_tensor().copy_(total)
    torch.cuda.synchronize()
    buf = gpu_render.createBuffer(BufferDesc(w, h, BufferFormat.FLOAT4))
    gpu_render.resolve(buf, 8)
    img = buf.map().copy()
    buf.destroy()
    st.setAs("render/b200/sampleOffset", 0)
    st.setAs("render/b200/sampleStride", 1)
    assert rel_rmse(img, full) < 1e-5


@pytest.mark.parametrize("tonemapper,gamma", [(1, 2.4), (2, 2.2), (3, 0.0), (0, 2.4)])
def test_postprocess_matches_reference_tonemappers(gpu_render, tonemapper, gamma):
    # Tonemappers.cu:17-135 on params.image after accumulation (OptixRender.cpp:1045-1049)
    s, st, _ = make_cornell(64, 64, 8)
    linear = _render(gpu_render, s, st, 64, 64, 8)
    st.setAs("render/pt/tonemapperType", tonemapper)
    st.setAs("render/post/gamma", gamma)
    img = _render(gpu_render, s, st, 64, 64, 8)
    ref = pyoracle.postprocess(linear, tonemapper, pyoracle.exposure(st), gamma)
    np.testing.assert_allclose(img[..., :3], ref[..., :3], rtol=2e-5, atol=1e-6)  # powf: CUDA vs glibc


def test_spp_greater_than_one_reproduces_reference_lerp_weights_q1(gpu_render):
    # spp = 3, sppTotal = 10 -> launches of 3, 3, 3, 1 samples, lerp weight 1/(subframe+1) (OptixRender.cu:60-78)
    s, st, _ = make_cornell(48, 48, 10)
    st.setAs("render/pt/spp", 3)
    img_g = _render(gpu_render, s, st, 48, 48, 4, batched=False)
    assert gpu_render.getSharedContext().mSubframeIndex == 10
    img_o, _, sub, _ = pyoracle.OracleScene(s).render(st, 48, 48, 4)
    assert sub == 10
    assert rel_rmse(img_g, img_o) <= 1e-5


def test_accumulation_disabled_returns_launch_mean(gpu_render):
    s, st, _ = make_cornell(48, 48, 16)
    st.setAs("render/pt/enableAcc", False)
    st.setAs("render/pt/spp", 2)
    a = _render(gpu_render, s, st, 48, 48, 1, batched=False)
    b = _render(gpu_render, s, st, 48, 48, 3, batched=False)  # every launch restarts at subframe 0
    assert np.array_equal(a, b)
    img_o, _, sub, _ = pyoracle.OracleScene(s).render(st, 48, 48, 1)
    assert sub == 0
    assert rel_rmse(a, img_o) <= 1e-5


def test_non_float4_output_formats(gpu_render):
    s, st, _ = make_cornell(32, 32, 4)
    r = gpu_render
    r.setScene(s)
    r.setSharedContext(SharedContext(mSettingsManager=st))
    r._last_settings = None
    r.reset_accumulation()
    f4 = r.createBuffer(BufferDesc(32, 32, BufferFormat.FLOAT4))
    r.render_iterations(f4, 4)
    ref = f4.map().copy()
    for fmt in (BufferFormat.FLOAT3, BufferFormat.UNSIGNED_BYTE4):
        b = r.createBuffer(BufferDesc(32, 32, fmt))
        r.resolve(b, 4)
        img = b.map().copy()
        if fmt == BufferFormat.FLOAT3:
            assert np.array_equal(img, ref[..., :3])
        else:
            assert img.shape == (32, 32, 4) and img.dtype == np.uint8
            np.testing.assert_allclose(img[..., :3], np.clip(ref[..., :3], 0, 1) * 255.0, atol=0.51)
        b.destroy()
    f4.destroy()


@pytest.mark.parametrize("debug,spp", [(2, 1), (3, 1), (3, 3)])
def test_aov_views_match_oracle_on_device(gpu_render, debug, spp):
    # diffuse / specular AOV views (OptixRender.cu:157-221) with their own per-pixel sample counters
    s, st = random_scene(seed=5)
    st.setAs("render/pt/sppTotal", 9)
    st.setAs("render/pt/spp", spp)
    st.setAs("render/pt/debug", debug)
    launches = 9 // spp
    img_g = _render(gpu_render, s, st, 40, 30, launches, batched=(spp == 1))
    assert gpu_render.getSharedContext().mSubframeIndex == 9
    img_o, _, sub, _ = pyoracle.OracleScene(s).render(st, 40, 30, launches)
    assert sub == 9
    assert img_o[..., :3].max() > 0.0
    assert rel_rmse(img_g, img_o) <= 1e-5


def test_fused_small_scene_kernel_is_bit_identical_to_wavefront(gpu_render):
    # SB_CFG_FUSED_SMALL: whole paths in one kernel, state in registers (kernels.cu: k_path_fused)
    from strelka_b200 import RenderFactory, RenderType

    s, st, _ = make_cornell(50, 38, 6)
    st.setAs("render/pt/depth", 5)
    gpu_render.reset_counters()  # the session-wide renderer has counted the earlier tests' rays
    img_w = _render(gpu_render, s, st, 50, 38, 6)
    cw = gpu_render.counters()
    fused = RenderFactory.createRender(RenderType.eCompute, fused_small=True)
    fused.init()
    try:
        img_f = _render(fused, s, st, 50, 38, 6)
        cf = fused.counters()
    finally:
        fused.destroy()
    assert cf["bvh_nodes_tri"] <= 64  # small enough to take the fused path
    assert np.array_equal(img_w, img_f)
    assert (cw["radiance_rays"], cw["shadow_rays"]) == (cf["radiance_rays"], cf["shadow_rays"])


def test_async_map_returns_the_frame_it_was_issued_for(gpu_render):
    # progressive display: copy of frame k overlaps the rendering of frame k+1
    s, st, _ = make_cornell(64, 64, 8)
    r = gpu_render
    r.setScene(s)
    r.setSharedContext(SharedContext(mSettingsManager=st))
    r._last_settings = None
    r.reset_accumulation()
    bufs = [r.createBuffer(BufferDesc(64, 64, BufferFormat.FLOAT4)) for _ in range(2)]  # round-robin like hdRunner
    r.render(bufs[0])
    bufs[0].map_async()
    r.render(bufs[1])  # issued while the copy of frame 0 may still be in flight
    frame0 = bufs[0].map_wait().copy()
    frame1 = bufs[1].map().copy()
    o = pyoracle.OracleScene(s)
    img1, acc, sub, _ = o.render(st, 64, 64, 1)
    img2, _, _, _ = o.render(st, 64, 64, 1, subframe=sub, accum=acc)
    assert rel_rmse(frame0, img1) <= 1e-5
    assert rel_rmse(frame1, img2) <= 1e-5
    assert not np.array_equal(frame0, frame1)
    for b in bufs:
        b.destroy()


def test_offset_ray_bit_exact_on_device(gpu_render):
    """offset_ray (closest_hit.cu:218-233, Ray Tracing Gems ch. 6): integer-ulp offsets away from the origin, float offsets
    near it -- integer and single float operations only, so device and oracle must agree bit for bit"""
    rng = np.random.default_rng(12)
    n = 100000
    p = np.concatenate([rng.uniform(-5, 5, (n // 2, 3)), rng.uniform(-1 / 16, 1 / 16, (n // 2, 3))]).astype(np.float32)
    p[:100] = 0.0
    p[100:200, 0] = np.float32(1.0 / 32.0)  # the switch-over magnitude itself
    nrm = rng.normal(size=(n, 3))
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    got = gpu_render.test_offset_ray(p, nrm)
    want = np.stack([pyoracle.offset_ray(p[i], nrm[i]) for i in range(0, n, 37)])
    assert np.array_equal(got[::37].view(np.uint32), want.view(np.uint32))


def test_raw_camera_matrices_reset_only_on_change(gpu_render):
    """a host that sets Params.clipToView / viewToWorld every frame keeps accumulating while they stand still
    (OptixRender.cpp:903-908 resets only when the matrices differ from the previous frame's)"""
    from strelka_b200.scenes import make_cornell

    w = h = 48
    s, st, _ = make_cornell(w, h, 16)
    r = gpu_render
    r.setScene(s)
    r.setSharedContext(SharedContext(mSettingsManager=st))
    r._last_settings = None
    buf = r.createBuffer(BufferDesc(w, h, BufferFormat.FLOAT4))
    r.render(buf)  # uploads scene, settings and the scene camera
    c2v, v2w = pyoracle.OracleScene(s).camera_matrices(w, h)
    r.set_camera_matrices(c2v, v2w)  # switching from the fov path: restart
    assert r._lib.sb_subframe_index(r._ctx) == 0
    for i in range(3):
        r.set_camera_matrices(c2v, v2w)  # unchanged: no restart
        r.render_raw(buf)
        assert r.getSharedContext().mSubframeIndex == i + 1
    a = buf.map().copy()
    ref, _, _, _ = pyoracle.OracleScene(s).render(st, w, h, 3)
    assert rel_rmse(a, ref) <= 1e-3  # the raw matrices are the ones the scene camera produces
    v2w2 = v2w.copy()
    v2w2[3] += 0.01  # move the eye
    r.set_camera_matrices(c2v, v2w2)
    assert r._lib.sb_subframe_index(r._ctx) == 0
    r._last_view = None  # hand the camera back to the scene for the tests that follow
    buf.destroy()


def test_buffer_resize_restarts_accumulation_like_reference(gpu_render):
    """the viewer resizes the output buffer in place (OptixBuffer.cpp:20-35); the next render() sees a new resolution,
    reallocates the accumulation buffers and restarts at subframe 0 (OptixRender.cpp:827-872, 910-934).  The resized
    frame must equal a fresh render at that size, bit for bit, and going back to the first size must do the same."""
    s, st, _ = make_cornell(96, 64, 8)
    fresh_small = _render(gpu_render, s, st, 64, 48, 3, batched=False)
    fresh_large = _render(gpu_render, s, st, 96, 64, 2, batched=False)
    r = gpu_render
    r.setScene(s)
    r.setSharedContext(SharedContext(mSettingsManager=st))
    r._last_settings = None
    r.reset_accumulation()
    buf = r.createBuffer(BufferDesc(64, 48, BufferFormat.FLOAT4))
    for _ in range(3):
        r.render(buf)
    assert r.getSharedContext().mSubframeIndex == 3
    assert np.array_equal(buf.map().copy(), fresh_small)
    buf.unmap()
    buf.resize(96, 64)
    assert buf.map().shape[:2] == (64, 96) and not buf.map().any()  # a resized buffer starts cleared
    buf.unmap()
    for i in range(2):
        r.render(buf)
        assert r.getSharedContext().mSubframeIndex == i + 1  # restarted
    assert np.array_equal(buf.map().copy(), fresh_large)
    buf.unmap()
    buf.resize(64, 48)
    for _ in range(3):
        r.render(buf)
    assert np.array_equal(buf.map().copy(), fresh_small)
    buf.unmap()
    buf.destroy()
