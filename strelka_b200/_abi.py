"""ctypes view of include/sb/sb_api.h (the C ABI) and numpy dtypes of its POD arrays.

The layouts here are checked against sizeof() values exported by the library
(tests/test_abi.py) so that a drift between the header and this file fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libstrelka_b200.so"


class SbError(RuntimeError):
    """Raised for every non-SB_OK result of the C ABI (the reference assert(0)s instead)."""


def library_path() -> str:
    # STRELKA_B200_LIB: alternative build of the same library (kernel-variant experiments); never a fallback
    return os.environ.get("STRELKA_B200_LIB") or os.path.join(_HERE, _LIB_NAME)


# ---- numpy dtypes of the scene arrays (byte-identical to the structs in sb_api.h) -------------
VERTEX_DTYPE = np.dtype(
    [("pos", "<f4", (3,)), ("tangent", "<u4"), ("normal", "<u4"), ("uv", "<u4"), ("pad0", "<f4"), ("pad1", "<f4")]
)
MESH_DTYPE = np.dtype([("index", "<u4"), ("count", "<u4"), ("vb_offset", "<u4"), ("vertex_count", "<u4")])
CURVE_DTYPE = np.dtype(
    [
        ("vertex_counts_start", "<u4"),
        ("vertex_counts_count", "<u4"),
        ("points_start", "<u4"),
        ("points_count", "<u4"),
        ("widths_start", "<u4"),
        ("widths_count", "<u4"),
    ]
)
INSTANCE_DTYPE = np.dtype(
    [("transform", "<f4", (16,)), ("type", "<u4"), ("geom_id", "<u4"), ("material_id", "<u4"), ("light_id", "<u4")]
)
LIGHT_DTYPE = np.dtype(
    [
        ("points", "<f4", (4, 4)),
        ("color", "<f4", (4,)),
        ("normal", "<f4", (4,)),
        ("type", "<i4"),
        ("half_angle", "<f4"),
        ("pad0", "<f4"),
        ("pad1", "<f4"),
    ]
)
MATERIAL_DTYPE = np.dtype(
    [
        ("model", "<u4"),
        ("base_color", "<f4", (3,)),
        ("roughness", "<f4"),
        ("metallic", "<f4"),
        ("ior", "<f4"),
        ("opacity", "<f4"),
        ("clearcoat", "<f4"),
        ("clearcoat_roughness", "<f4"),
        ("specular_color", "<f4", (3,)),
        ("use_specular_workflow", "<u4"),
        ("hair_absorption", "<f4", (3,)),
        ("hair_roughness_lon", "<f4"),
        ("hair_roughness_azi", "<f4"),
        ("hair_cuticle_angle", "<f4"),
        ("diffuse_texture", "<u4"),
        ("normal_texture", "<u4"),
        ("pad", "<f4", (2,)),
    ]
)
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4"), ("instance", "<u4"), ("kind", "<u4")])

assert VERTEX_DTYPE.itemsize == 32 and MESH_DTYPE.itemsize == 16 and CURVE_DTYPE.itemsize == 24
assert INSTANCE_DTYPE.itemsize == 80 and LIGHT_DTYPE.itemsize == 112 and MATERIAL_DTYPE.itemsize == 96
assert HIT_DTYPE.itemsize == 24

SB_INSTANCE_MESH, SB_INSTANCE_LIGHT, SB_INSTANCE_CURVE = 0, 1, 2
SB_MATERIAL_DIFFUSE, SB_MATERIAL_USD_PREVIEW_SURFACE, SB_MATERIAL_HAIR = 0, 1, 2
SB_FORMAT_UNSIGNED_BYTE4, SB_FORMAT_FLOAT4, SB_FORMAT_FLOAT3 = 0, 1, 2
SB_CFG_TRAVERSAL_STATS = 1
SB_CFG_STAGE_TIMERS = 2
SB_CFG_FUSED_SMALL = 4


class sb_scene_view(C.Structure):
    _fields_ = [
        ("vertices", C.c_void_p), ("num_vertices", C.c_uint64),
        ("indices", C.c_void_p), ("num_indices", C.c_uint64),
        ("meshes", C.c_void_p), ("num_meshes", C.c_uint32),
        ("curves", C.c_void_p), ("num_curves", C.c_uint32),
        ("curve_points", C.c_void_p), ("num_curve_points", C.c_uint64),
        ("curve_widths", C.c_void_p), ("num_curve_widths", C.c_uint64),
        ("curve_vertex_counts", C.c_void_p), ("num_curve_vertex_counts", C.c_uint64),
        ("instances", C.c_void_p), ("num_instances", C.c_uint32),
        ("lights", C.c_void_p), ("num_lights", C.c_uint32),
        ("materials", C.c_void_p), ("num_materials", C.c_uint32),
        ("textures", C.c_void_p), ("num_textures", C.c_uint32),
    ]


class sb_texture(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class sb_settings(C.Structure):
    _fields_ = [
        ("spp", C.c_uint32), ("spp_total", C.c_uint32), ("depth", C.c_uint32), ("enable_acc", C.c_uint32),
        ("rect_light_sampling_method", C.c_uint32), ("debug", C.c_uint32),
        ("shadow_ray_tmin", C.c_float), ("material_ray_tmin", C.c_float),
        ("tonemapper_type", C.c_uint32), ("gamma", C.c_float),
        ("film_iso", C.c_float), ("cm2_factor", C.c_float), ("f_stop", C.c_float), ("shutter_speed", C.c_float),
        ("sample_offset", C.c_uint32), ("sample_stride", C.c_uint32),
        ("reserved", C.c_uint32 * 4),
    ]


class sb_device_cfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_batch_paths", C.c_uint32), ("flags", C.c_uint32), ("curve_split", C.c_uint32)]


class sb_counters(C.Structure):
    _fields_ = [
        ("paths", C.c_uint64), ("radiance_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
        ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("segs_tested", C.c_uint64),
        ("stack_overflows", C.c_uint64),
        ("nodes_visited_shadow", C.c_uint64), ("tris_tested_shadow", C.c_uint64), ("segs_tested_shadow", C.c_uint64),
        ("num_triangles", C.c_uint64), ("num_segments", C.c_uint64),
        ("bvh_nodes_tri", C.c_uint64), ("bvh_nodes_curve", C.c_uint64),
        ("build_ms", C.c_double), ("render_ms", C.c_double),
        ("kernel_launches", C.c_uint64),
        ("stage_ms", C.c_double * 8), ("stage_launches", C.c_uint64 * 8),
        ("bvh_depth_tri", C.c_uint64), ("bvh_depth_curve", C.c_uint64),
        ("exchange_ms", C.c_double), ("exchange_nvls", C.c_uint64),
    ]

    STAGES = ("raygen", "extend", "shade", "shadow", "accumulate", "resolve", "path_fused", "reserved")

    def as_dict(self):
        d = {}
        for n, _ in self._fields_:
            v = getattr(self, n)
            d[n] = list(v) if hasattr(v, "__len__") else v
        return d


# Every symbol include/sb/sb_api.h declares (tests/test_abi.py checks they are all exported).
ABI_SYMBOLS = [
    "sb_abi_version", "sb_abi_struct_size", "sb_settings_default", "sb_create", "sb_set_stream", "sb_destroy", "sb_last_error", "sb_set_scene", "sb_set_camera",
    "sb_set_camera_matrices", "sb_set_settings", "sb_reset_accumulation", "sb_subframe_index",
    "sb_buffer_create", "sb_buffer_destroy", "sb_buffer_resize", "sb_buffer_map", "sb_buffer_unmap", "sb_buffer_map_async", "sb_buffer_map_wait",
    "sb_buffer_host_ptr", "sb_buffer_host_size", "sb_buffer_device_ptr", "sb_buffer_width", "sb_buffer_height",
    "sb_render", "sb_render_iterations", "sb_synchronize", "sb_accum_device_ptr", "sb_resolve",
    "sb_comm_get_unique_id", "sb_comm_init", "sb_comm_destroy", "sb_comm_world", "sb_comm_exchange_path", "sb_render_sharded",
    "sb_get_counters", "sb_reset_counters", "sb_test_sampler", "sb_test_light_sample", "sb_test_trace", "sb_test_bsdf", "sb_test_texture", "sb_test_offset_ray",
]
SB_COMM_ID_BYTES = 128
SB_API_VERSION = 2

_lib = None


def load_library() -> C.CDLL:
    """Load the CUDA backend.  Fails loudly when it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise SbError(
            f"{path} is missing: build the CUDA backend first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C strelka_b200/csrc). There is no CPU fallback."
        )
    lib = C.CDLL(path)
    vp, u32, u64, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float
    P = C.POINTER
    sig = {
        "sb_abi_version": (u32, []),
        "sb_abi_struct_size": (u32, [u32]),
        "sb_settings_default": (None, [P(sb_settings)]),
        "sb_create": (C.c_int, [P(sb_device_cfg), P(vp)]),
        "sb_set_stream": (C.c_int, [vp, vp]),
        "sb_destroy": (None, [vp]),
        "sb_last_error": (C.c_char_p, [vp]),
        "sb_set_scene": (C.c_int, [vp, P(sb_scene_view)]),
        "sb_set_camera": (C.c_int, [vp, P(f32), f32]),
        "sb_set_camera_matrices": (C.c_int, [vp, P(f32), P(f32)]),
        "sb_set_settings": (C.c_int, [vp, P(sb_settings)]),
        "sb_reset_accumulation": (C.c_int, [vp]),
        "sb_subframe_index": (u32, [vp]),
        "sb_buffer_create": (C.c_int, [vp, u32, u32, u32, P(vp)]),
        "sb_buffer_destroy": (None, [vp]),
        "sb_buffer_resize": (C.c_int, [vp, u32, u32]),
        "sb_buffer_map": (C.c_int, [vp, P(vp)]),
        "sb_buffer_map_async": (C.c_int, [vp]),
        "sb_buffer_map_wait": (C.c_int, [vp, P(vp)]),
        "sb_buffer_unmap": (C.c_int, [vp]),
        "sb_buffer_host_ptr": (vp, [vp]),
        "sb_buffer_host_size": (C.c_size_t, [vp]),
        "sb_buffer_device_ptr": (vp, [vp]),
        "sb_buffer_width": (u32, [vp]),
        "sb_buffer_height": (u32, [vp]),
        "sb_render": (C.c_int, [vp, vp]),
        "sb_render_iterations": (C.c_int, [vp, vp, u32]),
        "sb_synchronize": (C.c_int, [vp]),
        "sb_accum_device_ptr": (vp, [vp, P(u64)]),
        "sb_resolve": (C.c_int, [vp, vp, u32]),
        "sb_comm_get_unique_id": (C.c_int, [vp]),
        "sb_comm_init": (C.c_int, [vp, vp, u32, u32]),
        "sb_comm_destroy": (C.c_int, [vp]),
        "sb_comm_world": (u32, [vp]),
        "sb_comm_exchange_path": (C.c_char_p, [vp]),
        "sb_render_sharded": (C.c_int, [vp, vp, u32]),
        "sb_get_counters": (C.c_int, [vp, P(sb_counters)]),
        "sb_reset_counters": (C.c_int, [vp]),
        "sb_test_sampler": (C.c_int, [vp, u32, vp, vp, vp, vp, vp, vp, vp]),
        "sb_test_light_sample": (C.c_int, [vp, u32, vp, vp, vp, u32, vp]),
        "sb_test_trace": (C.c_int, [vp, u32, vp, u32, vp]),
        "sb_test_bsdf": (C.c_int, [vp, vp, u32, vp, vp]),
        "sb_test_texture": (C.c_int, [vp, u32, u32, vp, vp]),
        "sb_test_offset_ray": (C.c_int, [vp, u32, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    # layout handshake: a stale library (or a drifted mirror in this file) must not be driven
    mirrors = [C.sizeof(sb_settings), C.sizeof(sb_device_cfg), C.sizeof(sb_counters), C.sizeof(sb_scene_view), MATERIAL_DTYPE.itemsize,
               LIGHT_DTYPE.itemsize, INSTANCE_DTYPE.itemsize, VERTEX_DTYPE.itemsize, HIT_DTYPE.itemsize, MESH_DTYPE.itemsize,
               CURVE_DTYPE.itemsize, C.sizeof(sb_texture)]
    if lib.sb_abi_version() != SB_API_VERSION:
        raise SbError(f"{path} implements ABI version {lib.sb_abi_version()}, this package expects {SB_API_VERSION}: rebuild it")
    for which, size in enumerate(mirrors):
        if lib.sb_abi_struct_size(which) != size:
            raise SbError(f"ABI struct {which}: library says {lib.sb_abi_struct_size(which)} bytes, the Python mirror {size}")
    _lib = lib
    return lib


def np_ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data) if a is not None and a.size else C.c_void_p(0)
