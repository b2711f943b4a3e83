"""strelka_b200 -- B200-native (sm_100a) path-tracing backend for Strelka.

Host-side mirror of the reference's render interface (oka::Render / oka::Buffer / oka::Scene /
oka::Camera / SettingsManager) on top of the C ABI declared in include/sb/sb_api.h and implemented
by hand-written CUDA kernels in strelka_b200/csrc (built into strelka_b200/libstrelka_b200.so).

There is no CPU fallback: creating a Render without the CUDA library or without a GPU raises.
"""
from ._abi import (  # noqa: F401
    SbError,
    load_library,
    library_path,
    sb_settings,
    sb_counters,
    VERTEX_DTYPE,
    MESH_DTYPE,
    CURVE_DTYPE,
    INSTANCE_DTYPE,
    LIGHT_DTYPE,
    MATERIAL_DTYPE,
    HIT_DTYPE,
)
from .settings import SettingsManager, default_settings  # noqa: F401
from .camera import Camera  # noqa: F401
from .scene import Scene, UniformLightDesc, pack_normal, pack_uv, unpack_normal  # noqa: F401
from .render import Render, Buffer, BufferDesc, BufferFormat, RenderFactory, RenderType, SharedContext  # noqa: F401

__all__ = [
    "Render", "Buffer", "BufferDesc", "BufferFormat", "RenderFactory", "RenderType", "SharedContext",
    "Scene", "UniformLightDesc", "Camera", "SettingsManager", "default_settings", "SbError", "load_library",
]
