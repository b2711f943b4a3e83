"""TEST HARNESS ONLY: ctypes loader of tests/emul/_build/libemul.so (host instantiation of the
backend's device logic, see emul.cpp)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from strelka_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libemul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
        _lib = C.CDLL(_LIB)
        _lib.emul_scene_create.restype = C.c_void_p
        _lib.emul_scene_create.argtypes = [C.POINTER(_abi.sb_scene_view), C.c_uint32, C.c_char_p, C.c_int]
        _lib.emul_scene_destroy.argtypes = [C.c_void_p]
        _lib.emul_scene_info.argtypes = [C.c_void_p, C.c_void_p]
        _lib.emul_trace.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.emul_bvh_stats.argtypes = [C.c_void_p, C.c_void_p]
        _lib.emul_render.argtypes = [C.c_void_p, C.POINTER(_abi.sb_settings), C.c_void_p, C.c_float, C.c_uint32, C.c_uint32,
                                     C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        _lib.emul_sampler.argtypes = [C.c_uint32] + [C.c_void_p] * 7
        _lib.emul_light_sample.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        _lib.emul_curve_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.emul_camera.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        _lib.emul_octant_perm_table.argtypes = [C.c_void_p]
        _lib.emul_octant_perm_table.restype = C.c_int
    return _lib


class EmulScene:
    def __init__(self, scene, curve_split=8):
        self._scene = scene
        self._view = scene.view()
        err = C.create_string_buffer(512)
        self._h = lib().emul_scene_create(C.byref(self._view), curve_split, err, 512)
        if not self._h:
            raise RuntimeError(err.value.decode())

    def info(self):
        out = np.zeros(4, dtype=np.uint64)
        lib().emul_scene_info(self._h, out.ctypes.data)
        return {"triangles": int(out[0]), "segments": int(out[1]), "tri_nodes": int(out[2]), "seg_nodes": int(out[3])}

    def bvh_stats(self):
        out = np.zeros(7, dtype=np.uint64)
        lib().emul_bvh_stats(self._h, out.ctypes.data)
        keys = ("nodes", "inner_children", "leaf_slots", "prim_refs", "errors", "depth", "empty_slots")
        return dict(zip(keys, (int(v) for v in out)))

    def trace(self, rays, mode=0, with_stats=False):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hits = np.zeros(len(rays), dtype=_abi.HIT_DTYPE)
        stats = np.zeros(4, dtype=np.uint64)
        lib().emul_trace(self._h, len(rays), rays.ctypes.data, mode, hits.ctypes.data, stats.ctypes.data)
        return (hits, stats) if with_stats else hits

    def render(self, settings, width, height, samples, subframe=0, chunk_max=1, S=None, fused=False):
        st = settings.to_sb_settings() if hasattr(settings, "to_sb_settings") else settings
        cam = self._scene.getCamera(0)
        cam.updateViewMatrix()
        view = cam.view_glm()
        if S is None:
            S = np.zeros((height, width, 4), dtype=np.float32)
        image = np.zeros((height, width, 4), dtype=np.float32)
        counters = np.zeros(3, dtype=np.uint64)
        lib().emul_render(self._h, C.byref(st), view.ctypes.data, C.c_float(cam.fov), width, height, subframe, samples, chunk_max,
                          S.ctypes.data, image.ctypes.data, counters.ctypes.data, 1 if fused else 0)
        return image, S, {"paths": int(counters[0]), "radiance_rays": int(counters[1]), "shadow_rays": int(counters[2])}

    def close(self):
        if self._h:
            lib().emul_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sampler(x, y, sample, max_samples, depth, dim):
    arrs = [np.ascontiguousarray(a, dtype=np.uint32) for a in (x, y, sample, max_samples, depth, dim)]
    out = np.zeros(len(arrs[0]), dtype=np.float32)
    lib().emul_sampler(len(out), *[a.ctypes.data for a in arrs], out.ctypes.data)
    return out


def light_sample(lights, hit_points, u, method):
    lights = np.ascontiguousarray(lights, dtype=_abi.LIGHT_DTYPE)
    hp = np.ascontiguousarray(hit_points, dtype=np.float32)
    uu = np.ascontiguousarray(u, dtype=np.float32)
    out = np.zeros((len(lights), 12), dtype=np.float32)
    lib().emul_light_sample(len(lights), lights.ctypes.data, hp.ctypes.data, uu.ctypes.data, method, out.ctypes.data)
    return out


def curve_intersect(q, o, d, tmin=0.0, tmax=1e16):
    q = np.ascontiguousarray(q, dtype=np.float32)
    ray = np.array([*o, *d, tmin, tmax], dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    lib().emul_curve_intersect(q.ctypes.data, ray.ctypes.data, out.ctypes.data)
    return bool(out[0]), float(out[1]), float(out[2])


def camera(view, fov, aspect):
    view = np.ascontiguousarray(view, dtype=np.float32)
    c2v = np.zeros(16, dtype=np.float32)
    v2w = np.zeros(16, dtype=np.float32)
    lib().emul_camera(view.ctypes.data, C.c_float(fov), C.c_float(aspect), c2v.ctypes.data, v2w.ctypes.data)
    return c2v, v2w


def octant_perm_table():
    """(mismatches between the device's constexpr table and the host's bit-by-bit permutation, the 8 x 256 table)"""
    t = np.zeros(2048, np.uint8)
    bad = lib().emul_octant_perm_table(t.ctypes.data)
    return bad, t.reshape(8, 256)
