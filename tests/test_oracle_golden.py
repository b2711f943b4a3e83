"""The oracle against the reference's own headers (golden vectors from oracle/ref_crosscheck.cpp).

CPU only.  Bit-exact for the integer sampler; <= 1 ulp (in practice 0) for lights / tonemap / curve
normal math, where the only freedom is the evaluation order of dot/cross (see oracle/vec.h)."""
import numpy as np
import pytest

from conftest import bits_to_f32, ulp_diff
from oracle import pyoracle
from strelka_b200 import _abi


def test_sampler_integer_kats(golden):
    # SURVEY.md 8(c) KATs: Morton, murmur fmix, hash_combine, Laine-Karras, nested scramble, sobol
    assert list(pyoracle.sampler_ints()) == golden["sampler_ints"]
    assert golden["sampler_ints"] == [39, 917503, 33005907, 2140249156, 2160488919, 3815546535, 3758096384]


def test_sobol_direction_numbers_regenerated(golden):
    # our Joe-Kuo regeneration == the reference's literal sb_matrix[5][32]
    assert list(pyoracle.sobol_table()) == golden["sobol_matrix"]


def test_sampler_bit_exact(golden):
    i = np.array(golden["sampler_in"], dtype=np.uint32).reshape(-1, 6)
    out = pyoracle.sampler(i[:, 0], i[:, 1], i[:, 2], i[:, 3], i[:, 4], i[:, 5])
    assert np.array_equal(out.view(np.uint32), np.array(golden["sampler_out"], dtype=np.uint32))
    # the frozen decimal KAT of SURVEY.md 8(c), first row
    np.testing.assert_allclose(out[:5], [0.399413049, 0.479063481, 0.516170323, 0.268093169, 0.105077565], rtol=0, atol=1e-9)


def test_sampler_dimension_aliasing_quirk_q2():
    # eBSDF0 == ePixelX and eRussianRoulette == eLightPointY at every depth
    n = 64
    rng = np.random.default_rng(0)
    x, y, s = rng.integers(0, 1024, n), rng.integers(0, 1024, n), rng.integers(0, 256, n)
    mx = np.full(n, 256)
    for depth in range(4):
        d = np.full(n, depth)
        assert np.array_equal(pyoracle.sampler(x, y, s, mx, d, np.full(n, 5)), pyoracle.sampler(x, y, s, mx, d, np.full(n, 0)))
        assert np.array_equal(pyoracle.sampler(x, y, s, mx, d, np.full(n, 9)), pyoracle.sampler(x, y, s, mx, d, np.full(n, 4)))


def test_sampler_index_wraps_quirk_q3():
    # C5: Morton(x,y)*4096 wraps for x or y >= 1024 -> pixel (1024,0) aliases pixel (0,0)
    a = pyoracle.sampler([1024], [0], [5], [4096], [0], [0])
    b = pyoracle.sampler([0], [0], [5], [4096], [0], [0])
    assert a[0] == b[0]


def _lights(golden):
    raw = np.array(golden["light_structs"], dtype=np.uint32).reshape(-1, 28)
    lights = np.zeros(len(raw), dtype=_abi.LIGHT_DTYPE)
    lights.view(np.uint32).reshape(-1, 28)[:] = raw
    hp = bits_to_f32(golden["light_hit_points"]).reshape(-1, 3)
    u = bits_to_f32(golden["light_u"]).reshape(-1, 2)
    return lights, hp, u


@pytest.mark.parametrize("key,ltype,method", [
    ("light_sample_rect_uniform", 0, 0),
    ("light_sample_rect_sphquad", 0, 1),
    ("light_sample_sphere", 2, 0),
    ("light_sample_distant", 3, 0),
])
def test_light_sampling_matches_reference(golden, key, ltype, method):
    lights, hp, u = _lights(golden)
    sel = lights["type"] == ltype
    ref = bits_to_f32(golden[key]).reshape(-1, 12)
    got = pyoracle.light_sample(lights[sel], hp[sel], u[sel], method)
    assert got.shape == ref.shape
    d = ulp_diff(got, ref)
    # sphquad chains acos/sin/cos (libm on both sides here -> identical); allow 1 ulp headroom
    assert d.max() <= 1, f"{key}: max ulp diff {d.max()}"


def test_light_pdf_and_normal_match_reference(golden):
    lights, hp, _ = _lights(golden)
    raw = bits_to_f32(golden["light_pdf_normal"]).reshape(-1, 7)
    got = pyoracle.light_pdf(lights, raw[:, :3], hp)
    assert ulp_diff(got, raw[:, 3:7]).max() <= 1


def test_mis_balance(golden):
    v = bits_to_f32(golden["mis_balance"]).reshape(-1, 3)
    got = np.array([pyoracle.mis_balance(a, b) for a, b, _ in v], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), v[:, 2].copy().view(np.uint32))
    assert abs(pyoracle.mis_balance(0.5, 0.25) - 0.666667) < 1e-6


def test_tonemap_and_accumulate(golden):
    raw = np.array(golden["tonemap"], dtype=np.uint32).reshape(-1, 17)
    f = raw.view(np.float32)
    for row_u, row in zip(raw, f):
        c, c2, ev, sub = row[0:3], row[3:6], row[6], int(row_u[7])
        e = np.array([ev, ev, ev], dtype=np.float32)
        t = pyoracle.tonemap(0, c, c, e)
        assert np.array_equal(t.view(np.uint32), row_u[8:11])
        it = pyoracle.tonemap(1, t, t, e)
        assert np.array_equal(it.view(np.uint32), row_u[11:14])
        acc = pyoracle.tonemap(2, c, c2, e, sub)
        assert np.array_equal(acc.view(np.uint32), row_u[14:17])
    np.testing.assert_allclose(pyoracle.tonemap(0, [1, 2, 3], [0, 0, 0], [0.0625] * 3), [0.0588235, 0.111111, 0.157895], rtol=1e-5)


def test_curve_math(golden):
    raw = bits_to_f32(golden["curve"]).reshape(-1, 37)
    worst = 0
    for row in raw:
        q, u, ps, ref = row[:16], row[16], row[17:20], row[20:37]
        got = pyoracle.curve_eval(q, u, ps)
        worst = max(worst, int(ulp_diff(got, ref).max()))
    # position/velocity are exact; the normal goes through dot() whose evaluation order we fix
    # differently from the host-compiled reference -> a few ulp on cancelling sums
    assert worst <= 64, worst
    # SURVEY.md 8(c) KAT
    q = np.array([0, 0, 0, .1, 1, 0, 0, .1, 2, 1, 0, .1, 3, 1, 0, .05], dtype=np.float32)
    out = pyoracle.curve_eval(q, 0.5, [1.5, 0.6, 0.0])
    np.testing.assert_allclose(out[0:4], [1.5, 0.5, 0, 0.0989583], atol=1e-6)
    np.testing.assert_allclose(out[8:11], [0.8, 0.6, 0], atol=1e-6)
    np.testing.assert_allclose(out[11:14], [-0.595993, 0.80299, 0], atol=1e-5)
