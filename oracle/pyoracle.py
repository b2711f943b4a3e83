"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes loader for oracle/_build/liboracle.so (the CPU restatement of the reference's path).  Imported
only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs -- never by
the product package strelka_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from strelka_b200 import _abi  # POD layouts of include/sb/sb_api.h only

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_build/liboracle.so"])
    return _LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_scene_create.restype = C.c_void_p
        _lib.orc_scene_create.argtypes = [C.POINTER(_abi.sb_scene_view), C.c_uint32]
        _lib.orc_scene_destroy.argtypes = [C.c_void_p]
        _lib.orc_scene_info.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_render.restype = C.c_uint32
        _lib.orc_render.argtypes = [C.c_void_p, C.POINTER(_abi.sb_settings), C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                    C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib.orc_path_radiance.argtypes = [C.c_void_p, C.POINTER(_abi.sb_settings), C.c_void_p, C.c_void_p, C.c_uint32,
                                           C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib.orc_trace.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        _lib.orc_sampler.argtypes = [C.c_uint32] + [C.c_void_p] * 7
        _lib.orc_light_sample.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        _lib.orc_light_pdf.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_mis_balance.restype = C.c_float
        _lib.orc_mis_balance.argtypes = [C.c_float, C.c_float]
        _lib.orc_tonemap.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        _lib.orc_curve_eval.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        _lib.orc_curve_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_curve_intersect_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_offset_ray.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_texture_lookup.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.orc_bsdf_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.orc_camera_matrices.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        _lib.orc_postprocess.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_float]
        _lib.orc_exposure.argtypes = [C.POINTER(_abi.sb_settings), C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data


class OracleScene:
    """CPU scene (world-space BVH2s) built from a strelka_b200.Scene."""

    def __init__(self, scene, curve_split=8):
        self._scene = scene
        self._view = scene.view()
        self._h = lib().orc_scene_create(C.byref(self._view), curve_split)

    def info(self) -> dict:
        out = np.zeros(4, dtype=np.uint64)
        lib().orc_scene_info(self._h, _p(out))
        return {"triangles": int(out[0]), "segments": int(out[1]), "tri_nodes": int(out[2]), "seg_nodes": int(out[3])}

    def camera_matrices(self, width: int, height: int):
        cam = self._scene.getCamera(0)
        cam.updateViewMatrix()
        view = cam.view_glm()
        c2v = np.zeros(16, dtype=np.float32)
        v2w = np.zeros(16, dtype=np.float32)
        lib().orc_camera_matrices(_p(view), C.c_float(cam.fov), C.c_float(width / float(height)), _p(c2v), _p(v2w))
        return c2v, v2w

    def render(self, settings, width: int, height: int, launches: int, subframe: int = 0, accum=None, threads: int = 0, aov=None):
        """Emulate `launches` reference render() calls.  Returns (image, accum, new_subframe, counters)."""
        st = settings.to_sb_settings() if hasattr(settings, "to_sb_settings") else settings
        c2v, v2w = self.camera_matrices(width, height)
        if accum is None:
            accum = np.zeros((height, width, 4), dtype=np.float32)
        image = np.zeros((height, width, 4), dtype=np.float32)
        counters = np.zeros(3, dtype=np.uint64)
        sub = lib().orc_render(self._h, C.byref(st), _p(c2v), _p(v2w), width, height, subframe, launches, _p(accum), _p(image),
                               _p(counters), threads, _p(aov) if aov is not None else None)
        return image, accum, sub, {"paths": int(counters[0]), "radiance_rays": int(counters[1]), "shadow_rays": int(counters[2])}

    def path_radiance(self, settings, width, height, xs, ys, samples, threads: int = 0, counters: dict | None = None) -> np.ndarray:
        """Radiance of individual paths (x, y, sample index) at the given resolution, on `threads` host threads
        (0 = all).  `counters` (optional dict) receives paths / radiance_rays / shadow_rays."""
        st = settings.to_sb_settings() if hasattr(settings, "to_sb_settings") else settings
        c2v, v2w = self.camera_matrices(width, height)
        xs, ys, samples = [np.ascontiguousarray(a, dtype=np.uint32) for a in (xs, ys, samples)]
        out = np.zeros((len(xs), 3), dtype=np.float32)
        cnt = np.zeros(3, dtype=np.uint64)
        lib().orc_path_radiance(self._h, C.byref(st), _p(c2v), _p(v2w), width, height, len(xs), _p(xs), _p(ys), _p(samples), _p(out),
                                threads, _p(cnt))
        if counters is not None:
            counters.update(paths=int(cnt[0]), radiance_rays=int(cnt[1]), shadow_rays=int(cnt[2]))
        return out

    def texture_lookup(self, index: int, uv) -> np.ndarray:
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        out = np.zeros((len(uv), 4), dtype=np.float32)
        lib().orc_texture_lookup(self._h, index, len(uv), _p(uv), _p(out))
        return out

    def trace(self, rays, mode: int = 0) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hits = np.zeros(len(rays), dtype=_abi.HIT_DTYPE)
        lib().orc_trace(self._h, len(rays), _p(rays), mode, _p(hits))
        return hits

    def close(self):
        if self._h:
            lib().orc_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sampler(x, y, sample, max_samples, depth, dim) -> np.ndarray:
    arrs = [np.ascontiguousarray(a, dtype=np.uint32) for a in (x, y, sample, max_samples, depth, dim)]
    out = np.zeros(len(arrs[0]), dtype=np.float32)
    lib().orc_sampler(len(out), *[_p(a) for a in arrs], _p(out))
    return out


def sampler_ints() -> np.ndarray:
    out = np.zeros(7, dtype=np.uint32)
    lib().orc_sampler_ints(_p(out))
    return out


def sobol_table() -> np.ndarray:
    out = np.zeros(160, dtype=np.uint32)
    lib().orc_sobol_table(_p(out))
    return out


def light_sample(lights, hit_points, u, method) -> np.ndarray:
    lights = np.ascontiguousarray(lights, dtype=_abi.LIGHT_DTYPE)
    hp = np.ascontiguousarray(hit_points, dtype=np.float32)
    uu = np.ascontiguousarray(u, dtype=np.float32)
    out = np.zeros((len(lights), 12), dtype=np.float32)
    lib().orc_light_sample(len(lights), _p(lights), _p(hp), _p(uu), method, _p(out))
    return out


def light_pdf(lights, light_hits, surface_hits) -> np.ndarray:
    lights = np.ascontiguousarray(lights, dtype=_abi.LIGHT_DTYPE)
    lh = np.ascontiguousarray(light_hits, dtype=np.float32)
    sh = np.ascontiguousarray(surface_hits, dtype=np.float32)
    out = np.zeros((len(lights), 4), dtype=np.float32)
    lib().orc_light_pdf(len(lights), _p(lights), _p(lh), _p(sh), _p(out))
    return out


def mis_balance(a, b) -> float:
    return float(lib().orc_mis_balance(C.c_float(a), C.c_float(b)))


def tonemap(mode, c, c2, exposure, subframe=0) -> np.ndarray:
    c = np.ascontiguousarray(c, dtype=np.float32)
    c2 = np.ascontiguousarray(c2, dtype=np.float32)
    e = np.ascontiguousarray(exposure, dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    lib().orc_tonemap(mode, _p(c), _p(c2), _p(e), subframe, _p(out))
    return out


def curve_eval(q, u, ps) -> np.ndarray:
    q = np.ascontiguousarray(q, dtype=np.float32)
    ps = np.ascontiguousarray(ps, dtype=np.float32)
    out = np.zeros(17, dtype=np.float32)
    lib().orc_curve_eval(_p(q), C.c_float(u), _p(ps), _p(out))
    return out


def curve_intersect(q, o, d, tmin=0.0, tmax=1e16, f32=False):
    """Ray vs round cubic B-spline segment: the double bracketing solver, or (f32=True) the float solver."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    ray = np.array([*o, *d, tmin, tmax], dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    (lib().orc_curve_intersect_f32 if f32 else lib().orc_curve_intersect)(_p(q), _p(ray), _p(out))
    return bool(out[0]), float(out[1]), float(out[2])


def offset_ray(p, n) -> np.ndarray:
    p = np.ascontiguousarray(p, dtype=np.float32)
    n = np.ascontiguousarray(n, dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    lib().orc_offset_ray(_p(p), _p(n), _p(out))
    return out


def bsdf_batch(material, n, ng, tangent, k1, xi, k2) -> tuple[np.ndarray, np.ndarray]:
    """BSDF protocol of closest_hit.cu:474-545 for N items: sample(k1, xi) and evaluate(k1, k2).
    Inputs broadcast to (N, 3) / (N, 4).  Returns (sample (N, 8): k2, bsdf_over_pdf, pdf, event; eval (N, 7): diffuse, glossy, pdf)."""
    m = np.ascontiguousarray(material, dtype=_abi.MATERIAL_DTYPE)
    inp = pack_bsdf_inputs(n, ng, tangent, k1, xi, k2)
    out = np.zeros((len(inp), 15), dtype=np.float32)
    lib().orc_bsdf_batch(_p(m), len(inp), _p(inp), _p(out))
    return out[:, :8], out[:, 8:]


def pack_bsdf_inputs(n, ng, tangent, k1, xi, k2) -> np.ndarray:
    cols = [np.atleast_2d(np.asarray(a, dtype=np.float32)) for a in (n, ng, tangent, k1, xi, k2)]
    count = max(len(c) for c in cols)
    cols = [np.broadcast_to(c, (count, c.shape[1])) for c in cols]
    return np.ascontiguousarray(np.concatenate(cols, axis=1), dtype=np.float32)


def bsdf(material, n, ng, k1, xi, k2, tangent=(1.0, 0.0, 0.0)):
    s, e = bsdf_batch(material, n, ng, tangent, k1, xi, k2)
    return s[0].copy(), e[0].copy()


def postprocess(image, tonemapper_type, exposure, gamma) -> np.ndarray:
    img = np.ascontiguousarray(image, dtype=np.float32).copy()
    e = np.ascontiguousarray(exposure, dtype=np.float32)
    lib().orc_postprocess(_p(img), img.size // 4, tonemapper_type, _p(e), C.c_float(gamma))
    return img


def exposure(settings) -> np.ndarray:
    st = settings.to_sb_settings() if hasattr(settings, "to_sb_settings") else settings
    out = np.zeros(3, dtype=np.float32)
    lib().orc_exposure(C.byref(st), _p(out))
    return out
