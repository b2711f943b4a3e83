"""Host-side mirror of the reference's render interface on top of the C ABI.

Reference: include/render/render.h:9-63 (Render, RenderFactory, RenderType), include/render/buffer.h:
9-88 (Buffer, BufferDesc, BufferFormat), include/render/common.h:22-28 (SharedContext).  Names,
argument meaning and call order follow the reference so tests read like a Strelka client:

    render = RenderFactory.createRender(RenderType.eCompute)
    render.setScene(scene); render.setSharedContext(ctx); render.init()
    buf = render.createBuffer(BufferDesc(w, h, BufferFormat.FLOAT4))
    render.render(buf); buf.map(); pixels = buf.getHostPointer()
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass

import numpy as np

from . import _abi
from ._abi import SbError, sb_counters, sb_device_cfg
from .scene import Scene
from .settings import SettingsManager


class RenderType(enum.IntEnum):  # render.h:9-14
    eOptiX = 0
    eMetal = 1
    eCompute = 2


class BufferFormat(enum.IntEnum):  # buffer.h:9-14
    UNSIGNED_BYTE4 = 0
    FLOAT4 = 1
    FLOAT3 = 2


@dataclass
class BufferDesc:  # buffer.h:16-21
    width: int
    height: int
    format: BufferFormat = BufferFormat.FLOAT4


@dataclass
class SharedContext:  # common.h:22-28
    mFrameNumber: int = 0  # noqa: N815
    mSubframeIndex: int = 0  # noqa: N815
    mSettingsManager: SettingsManager | None = None  # noqa: N815
    mRender: "Render | None" = None  # noqa: N815


def _check(lib, ctx, rc, what):
    if rc != 0:
        msg = lib.sb_last_error(ctx)
        raise SbError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


class Buffer:
    """oka::Buffer (buffer.h:23-88) over sb_buffer."""

    _ELEM = {BufferFormat.FLOAT4: 16, BufferFormat.FLOAT3: 12, BufferFormat.UNSIGNED_BYTE4: 4}

    def __init__(self, render: "Render", desc: BufferDesc):
        self._render = render
        self._lib = render._lib
        self._format = BufferFormat(desc.format)
        h = C.c_void_p()
        _check(self._lib, render._ctx, self._lib.sb_buffer_create(render._ctx, desc.width, desc.height, int(desc.format), C.byref(h)),
               "sb_buffer_create")
        self._h = h

    def resize(self, width: int, height: int) -> None:
        _check(self._lib, self._render._ctx, self._lib.sb_buffer_resize(self._h, width, height), "sb_buffer_resize")

    def map(self):
        """Blocking device->host copy (OptixBuffer.cpp:37-43).  Returns the host array."""
        p = C.c_void_p()
        _check(self._lib, self._render._ctx, self._lib.sb_buffer_map(self._h, C.byref(p)), "sb_buffer_map")
        return self.getHostPointer()

    def map_async(self) -> None:
        """Enqueue the device->host copy behind the rendering issued so far and return at once (sb_buffer_map_async)."""
        _check(self._lib, self._render._ctx, self._lib.sb_buffer_map_async(self._h), "sb_buffer_map_async")

    def map_wait(self):
        """Wait for the copy of the last map_async() only; later render calls may still be running."""
        p = C.c_void_p()
        _check(self._lib, self._render._ctx, self._lib.sb_buffer_map_wait(self._h, C.byref(p)), "sb_buffer_map_wait")
        return self.getHostPointer()

    def unmap(self) -> None:
        _check(self._lib, self._render._ctx, self._lib.sb_buffer_unmap(self._h), "sb_buffer_unmap")

    def width(self) -> int:
        return self._lib.sb_buffer_width(self._h)

    def height(self) -> int:
        return self._lib.sb_buffer_height(self._h)

    def getFormat(self) -> BufferFormat:  # noqa: N802
        return self._format

    def getElementSize(self) -> int:  # noqa: N802
        return self._ELEM[self._format]

    def getHostDataSize(self) -> int:  # noqa: N802
        return self._lib.sb_buffer_host_size(self._h)

    def getHostPointer(self) -> np.ndarray:  # noqa: N802
        """Host mirror as a numpy view (h, w, channels)."""
        ptr = self._lib.sb_buffer_host_ptr(self._h)
        n = self.getHostDataSize()
        w, h = self.width(), self.height()
        if self._format == BufferFormat.UNSIGNED_BYTE4:
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,)).reshape(h, w, 4)
        ch = 4 if self._format == BufferFormat.FLOAT4 else 3
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n // 4,)).reshape(h, w, ch)

    def getNativePtr(self) -> int:  # noqa: N802
        return self._lib.sb_buffer_device_ptr(self._h)

    def destroy(self) -> None:
        if self._h:
            self._lib.sb_buffer_destroy(self._h)
            self._h = None


class Render:
    """oka::Render (render.h:19-56) implemented by the B200 backend (RenderType::eCompute)."""

    def __init__(self, device: int = 0, traversal_stats: bool = False, max_batch_paths: int = 0, stage_timers: bool = False,
                 curve_split: int = 0, fused_small: bool = False):
        self._lib = _abi.load_library()
        self._ctx = None
        self._device = device
        self._flags = ((_abi.SB_CFG_TRAVERSAL_STATS if traversal_stats else 0) | (_abi.SB_CFG_STAGE_TIMERS if stage_timers else 0)
                       | (_abi.SB_CFG_FUSED_SMALL if fused_small else 0))
        self._max_batch = max_batch_paths
        self._curve_split = curve_split
        self.mSharedCtx: SharedContext | None = None  # noqa: N815
        self.mScene: Scene | None = None  # noqa: N815
        self._scene_uploaded = False
        self._last_view = None
        self._last_fov = None
        self._last_settings = None

    # -- non-virtual setters of oka::Render (render.h:33-51)
    def setSharedContext(self, ctx: SharedContext) -> None:  # noqa: N802
        self.mSharedCtx = ctx

    def getSharedContext(self) -> SharedContext:  # noqa: N802
        return self.mSharedCtx

    def setScene(self, scene: Scene) -> None:  # noqa: N802
        self.mScene = scene
        self._scene_uploaded = False

    def getScene(self) -> Scene:  # noqa: N802
        return self.mScene

    # -- virtuals
    def init(self) -> None:
        cfg = sb_device_cfg(self._device, self._max_batch, self._flags, self._curve_split)
        h = C.c_void_p()
        rc = self._lib.sb_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            msg = self._lib.sb_last_error(None)
            raise SbError(f"sb_create failed ({rc}): {msg.decode() if msg else ''}")
        self._ctx = h

    def createBuffer(self, desc: BufferDesc) -> Buffer:  # noqa: N802
        return Buffer(self, desc)

    def getNativeDevicePtr(self):  # noqa: N802
        return None

    def _sync_inputs(self) -> None:
        lib, ctx = self._lib, self._ctx
        if not self._scene_uploaded:
            # frame 0 of OptiXRender::render (OptixRender.cpp:876-888)
            v = self.mScene.view()
            _check(lib, ctx, lib.sb_set_scene(ctx, C.byref(v)), "sb_set_scene")
            self._scene_uploaded = True
            self._last_view = None
        cam = self.mScene.getCamera(0)
        cam.updateViewMatrix()
        view = cam.view_glm()
        if self._last_view is None or not np.array_equal(view, self._last_view) or cam.fov != self._last_fov:
            _check(lib, ctx, lib.sb_set_camera(ctx, view.ctypes.data_as(C.POINTER(C.c_float)), float(cam.fov)), "sb_set_camera")
            self._last_view, self._last_fov = view.copy(), cam.fov
        st = self.mSharedCtx.mSettingsManager.to_sb_settings()
        raw = bytes(st)
        if raw != self._last_settings:
            _check(lib, ctx, lib.sb_set_settings(ctx, C.byref(st)), "sb_set_settings")
            self._last_settings = raw

    def render(self, output: Buffer) -> None:
        """OptiXRender::render(Buffer*) (OptixRender.cpp:874-1057): one launch of `spp` samples."""
        self._sync_inputs()
        _check(self._lib, self._ctx, self._lib.sb_render(self._ctx, output._h), "sb_render")
        self.mSharedCtx.mSubframeIndex = self._lib.sb_subframe_index(self._ctx)
        self.mSharedCtx.mFrameNumber += 1

    def render_iterations(self, output: Buffer, iterations: int) -> None:
        """`iterations` consecutive render() calls of the reference app loop, pipelined."""
        self._sync_inputs()
        _check(self._lib, self._ctx, self._lib.sb_render_iterations(self._ctx, output._h, iterations), "sb_render_iterations")
        self.mSharedCtx.mSubframeIndex = self._lib.sb_subframe_index(self._ctx)
        self.mSharedCtx.mFrameNumber += iterations

    def set_stream(self, cuda_stream: int | None) -> None:
        """Render on a caller-owned CUDA stream (e.g. torch.cuda.current_stream().cuda_stream; 0 is CUDA's
        legacy default stream).  None restores the context's private stream (SB_STREAM_PRIVATE)."""
        h = C.c_void_p(-1) if cuda_stream is None else C.c_void_p(cuda_stream)
        _check(self._lib, self._ctx, self._lib.sb_set_stream(self._ctx, h), "sb_set_stream")

    def set_camera_matrices(self, clip_to_view, view_to_world) -> None:
        """Raw Params.clipToView / viewToWorld (row-major float[16] each, OptixRender.cpp:953-954) instead of the camera of
        the scene; accumulation restarts only when the matrices differ from the previous call's (OptixRender.cpp:903-908).
        Use with sb_render through render_raw()."""
        a = np.ascontiguousarray(clip_to_view, dtype=np.float32).reshape(16)
        b = np.ascontiguousarray(view_to_world, dtype=np.float32).reshape(16)
        _check(self._lib, self._ctx, self._lib.sb_set_camera_matrices(self._ctx, a.ctypes.data_as(C.POINTER(C.c_float)),
                                                                      b.ctypes.data_as(C.POINTER(C.c_float))), "sb_set_camera_matrices")

    def render_raw(self, output: Buffer) -> None:
        """sb_render without re-sending the scene camera (for hosts that drive the raw matrices)"""
        _check(self._lib, self._ctx, self._lib.sb_render(self._ctx, output._h), "sb_render")
        self.mSharedCtx.mSubframeIndex = self._lib.sb_subframe_index(self._ctx)

    def synchronize(self) -> None:
        _check(self._lib, self._ctx, self._lib.sb_synchronize(self._ctx), "sb_synchronize")

    def reset_accumulation(self) -> None:
        _check(self._lib, self._ctx, self._lib.sb_reset_accumulation(self._ctx), "sb_reset_accumulation")
        self.mSharedCtx.mSubframeIndex = 0

    def counters(self) -> dict:
        c = sb_counters()
        _check(self._lib, self._ctx, self._lib.sb_get_counters(self._ctx, C.byref(c)), "sb_get_counters")
        return c.as_dict()

    def reset_counters(self) -> None:
        _check(self._lib, self._ctx, self._lib.sb_reset_counters(self._ctx), "sb_reset_counters")

    # -- multi-GPU plumbing
    def accum_device_ptr(self):
        n = C.c_uint64()
        p = self._lib.sb_accum_device_ptr(self._ctx, C.byref(n))
        return p, n.value

    def upload_scene_view(self, view) -> None:
        """sb_set_scene on an already flattened (e.g. pinned) sb_scene_view of self.mScene."""
        _check(self._lib, self._ctx, self._lib.sb_set_scene(self._ctx, C.byref(view)), "sb_set_scene")
        self._scene_uploaded = True

    def accum_tensor_nosync(self):
        return self.accum_tensor(sync=False)

    def accum_tensor(self, sync: bool = True):
        """Zero-copy torch view (float32, [h*w*4]) of the device accumulation buffer S = sum_k T(L_k): the
        quantity a multi-GPU harness all-reduces (torch.distributed / NCCL).  With sync=True the context's
        stream is synchronised first; pass sync=False when the context renders on torch's current stream
        (set_stream), where stream order already guarantees visibility."""
        import torch

        if sync:
            self.synchronize()
        ptr, n = self.accum_device_ptr()

        class _Wrap:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}

        return torch.as_tensor(_Wrap(), device=f"cuda:{self._device}")

    @staticmethod
    def _prefer_bundled_nccl() -> None:
        """A Python process that (later) imports torch must share ONE libnccl with it: torch's wheel links against its
        bundled copy (nvidia/nccl/lib/libnccl.so.2) and the dynamic loader reuses whatever object already carries that
        SONAME -- an older system libnccl loaded first breaks `import torch` (undefined ncclDevCommCreate).  So point
        the library's dlopen at the bundled copy when there is one (STRELKA_B200_NCCL, read by sb_comm_*)."""
        import importlib.util
        import os

        if os.environ.get("STRELKA_B200_NCCL"):
            return
        try:
            spec = importlib.util.find_spec("nvidia.nccl")
        except (ImportError, ValueError):
            spec = None
        for base in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["STRELKA_B200_NCCL"] = cand
                return

    def comm_unique_id(self) -> bytes:
        """ncclGetUniqueId through the ABI (rank 0); hand the bytes to the other ranks out of band."""
        self._prefer_bundled_nccl()
        buf = C.create_string_buffer(_abi.SB_COMM_ID_BYTES)
        _check(self._lib, None, self._lib.sb_comm_get_unique_id(buf), "sb_comm_get_unique_id")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int) -> None:
        """Join the NCCL group (collective).  The settings manager gets the matching sharding keys."""
        assert len(unique_id) == _abi.SB_COMM_ID_BYTES
        self._prefer_bundled_nccl()
        _check(self._lib, self._ctx, self._lib.sb_comm_init(self._ctx, unique_id, rank, world), "sb_comm_init")
        sm = self.mSharedCtx.mSettingsManager
        sm.setAs("render/b200/sampleOffset", int(rank))
        sm.setAs("render/b200/sampleStride", int(world))
        self._last_settings = None

    def comm_world(self) -> int:
        return int(self._lib.sb_comm_world(self._ctx)) if self._ctx else 1

    def comm_exchange_path(self) -> str:
        """how the group sums S: the fused NVLS kernel, or ncclAllReduce + resolve (and why)"""
        return (self._lib.sb_comm_exchange_path(self._ctx) or b"").decode()

    def comm_destroy(self) -> None:
        _check(self._lib, self._ctx, self._lib.sb_comm_destroy(self._ctx), "sb_comm_destroy")

    def render_sharded(self, output: Buffer, iterations: int) -> None:
        """This rank's share of `iterations` samples per rank, NCCL sum of S inside the library, global resolve."""
        self._sync_inputs()
        _check(self._lib, self._ctx, self._lib.sb_render_sharded(self._ctx, output._h, iterations), "sb_render_sharded")
        self.mSharedCtx.mSubframeIndex = self._lib.sb_subframe_index(self._ctx)
        self.mSharedCtx.mFrameNumber += iterations

    def resolve(self, output: Buffer, total_samples: int) -> None:
        _check(self._lib, self._ctx, self._lib.sb_resolve(self._ctx, output._h, total_samples), "sb_resolve")

    # -- test hooks
    def test_sampler(self, x, y, sample, max_samples, depth, dim) -> np.ndarray:
        arrs = [np.ascontiguousarray(a, dtype=np.uint32) for a in (x, y, sample, max_samples, depth, dim)]
        n = len(arrs[0])
        out = np.zeros(n, dtype=np.float32)
        _check(self._lib, self._ctx, self._lib.sb_test_sampler(self._ctx, n, *[a.ctypes.data for a in arrs], out.ctypes.data),
               "sb_test_sampler")
        return out

    def test_light_sample(self, lights, hit_points, u, method) -> np.ndarray:
        lights = np.ascontiguousarray(lights, dtype=_abi.LIGHT_DTYPE)
        hp = np.ascontiguousarray(hit_points, dtype=np.float32)
        uu = np.ascontiguousarray(u, dtype=np.float32)
        n = len(lights)
        out = np.zeros((n, 12), dtype=np.float32)
        _check(self._lib, self._ctx,
               self._lib.sb_test_light_sample(self._ctx, n, lights.ctypes.data, hp.ctypes.data, uu.ctypes.data, method, out.ctypes.data),
               "sb_test_light_sample")
        return out

    def test_offset_ray(self, p, normal) -> np.ndarray:
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normal, dtype=np.float32).reshape(-1, 3)
        out = np.zeros_like(p)
        _check(self._lib, self._ctx, self._lib.sb_test_offset_ray(self._ctx, len(p), p.ctypes.data, nrm.ctypes.data, out.ctypes.data), "sb_test_offset_ray")
        return out

    def test_texture(self, index: int, uv) -> np.ndarray:
        """Hardware-filtered lookups of texture `index` (0-based) of the current scene at (N, 2) st coordinates."""
        if not self._scene_uploaded:
            v = self.mScene.view()
            _check(self._lib, self._ctx, self._lib.sb_set_scene(self._ctx, C.byref(v)), "sb_set_scene")
            self._scene_uploaded = True
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        out = np.zeros((len(uv), 4), dtype=np.float32)
        _check(self._lib, self._ctx, self._lib.sb_test_texture(self._ctx, index, len(uv), uv.ctypes.data, out.ctypes.data), "sb_test_texture")
        return out

    def test_bsdf(self, material, packed_inputs) -> tuple[np.ndarray, np.ndarray]:
        """Device BSDF sample + evaluate on (N, 19) packed inputs (n, ng, tangent, k1, xi, k2); returns (sample (N, 8), eval (N, 7))."""
        m = np.ascontiguousarray(material, dtype=_abi.MATERIAL_DTYPE)
        inp = np.ascontiguousarray(packed_inputs, dtype=np.float32).reshape(-1, 19)
        out = np.zeros((len(inp), 15), dtype=np.float32)
        _check(self._lib, self._ctx, self._lib.sb_test_bsdf(self._ctx, m.ctypes.data, len(inp), inp.ctypes.data, out.ctypes.data), "sb_test_bsdf")
        return out[:, :8], out[:, 8:]

    def test_trace(self, rays, mode: int = 0) -> np.ndarray:
        if not self._scene_uploaded:
            v = self.mScene.view()
            _check(self._lib, self._ctx, self._lib.sb_set_scene(self._ctx, C.byref(v)), "sb_set_scene")
            self._scene_uploaded = True
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hits = np.zeros(len(rays), dtype=_abi.HIT_DTYPE)
        _check(self._lib, self._ctx, self._lib.sb_test_trace(self._ctx, len(rays), rays.ctypes.data, mode, hits.ctypes.data),
               "sb_test_trace")
        return hits

    def destroy(self) -> None:
        if self._ctx:
            self._lib.sb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class RenderFactory:  # render.h:58-63, render.cpp:10-35
    @staticmethod
    def createRender(type_: RenderType = RenderType.eCompute, **kw) -> Render | None:  # noqa: N802
        if type_ == RenderType.eCompute:
            return Render(**kw)
        return None  # "unsupported" -> nullptr, like render.cpp:17-18,25
