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


def render_sharded(render, buffer, iterations: int, group=None):
    """Render this rank's share of `iterations` samples per rank and resolve the global image on every rank.

    With an sb_comm group on the render (Render.comm_init) everything happens inside the library: NCCL sum of S into
    a separate buffer on the render's stream, resolve with the global sample count (sb_render_sharded).  Otherwise
    (torch.distributed group, e.g. the gloo/CPU tests or a host that already owns a process group) S is summed with
    torch on the current stream, resolved, and RESTORED to this rank's partial sum, so the call can be repeated and
    later progressive renders keep accumulating correctly.  Returns the samples this rank rendered."""
    before = render.getSharedContext().mSubframeIndex
    if render.comm_world() > 1:
        render.render_sharded(buffer, iterations)
        return render.getSharedContext().mSubframeIndex - before
    import torch.distributed as dist

    render.render_iterations(buffer, iterations)  # stops at this rank's local budget
    done = render.getSharedContext().mSubframeIndex - before
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        import torch

        s_probe = render.accum_tensor(sync=False)
        count = torch.tensor([float(render.getSharedContext().mSubframeIndex)], dtype=torch.float64, device=s_probe.device)
        dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
        n_total = int(count.item())  # samples accumulated over all ranks
        s_local = render.accum_tensor(sync=False)
        keep = s_local.clone()
        allreduce_accumulation(s_local, group)
        render.resolve(buffer, n_total)
        s_local.copy_(keep)
    return done
