"""Sample-index data parallelism over torch.distributed -- new in this backend, the
reference is single-GPU (OptixRender.cpp:163-189).

One process per GPU; every rank holds a full scene/BVH replica and renders the sample indices
rank, rank+world, ... of every pixel with the reference's sampler indices unchanged
(maxSampleCount = global sppTotal).  The only exchange is one sum all-reduce of the accumulation buffer
S = sum_k T(L_k) (float4 per pixel), after which every rank resolves T^-1(S / n_total).
"""
from __future__ import annotations


def shard_settings(settings, rank: int, world: int):
    """Write the sharding keys into a SettingsManager (no reference key: render/b200/*)."""
    settings.setAs("render/b200/sampleOffset", int(rank))
    settings.setAs("render/b200/sampleStride", int(world))
    return settings


def local_sample_count(spp_total: int, rank: int, world: int) -> int:
    """How many of the global sample indices [0, spp_total) belong to `rank` (offset + k*stride)."""
    if rank >= spp_total:
        return 0
    return (spp_total - rank + world - 1) // world


def allreduce_accumulation(tensor, group=None):
    """Sum-all-reduce the flat S tensor in place (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def render_sharded(render, buffer, spp_total: int, group=None):
    """Render this rank's share, all-reduce S over the render's stream and resolve the global image.

    `render` must have been put on torch's current stream (Render.set_stream) so that the collective is
    stream-ordered with the kernels.  Returns the number of samples this rank rendered."""
    import torch.distributed as dist

    before = render.getSharedContext().mSubframeIndex
    render.render_iterations(buffer, spp_total)  # stops at this rank's local budget
    done = render.getSharedContext().mSubframeIndex - before
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        allreduce_accumulation(render.accum_tensor(sync=False), group)
        render.resolve(buffer, spp_total)
    return done
