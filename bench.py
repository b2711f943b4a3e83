#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 path-tracing backend (contract: see README / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): synthetic Cornell box, diffuse + one rect light,
1024x1024, 256 spp (1 spp per iteration), max depth 4.  One STEP = one complete render of that
configuration (256 iterations of the hot path over all pixels).

Metric: Mrays/s = (radiance rays + shadow rays actually traced, device counters) / time / 1e6, whole job.
  value : device time (CUDA events on the render stream), scene + BVH already resident in HBM
  e2e   : the same metric through the public API with HOST buffers -- scene upload from pinned host
          memory + on-device BVH build + render + device->host read of the image, wall clock
N > 1 (torchrun): one process per GPU, each holds a scene/BVH replica and renders a disjoint stride of
sample indices (rank r: r, r+N, ...), then one NCCL all-reduce of the accumulation buffer S (float4 per
pixel) + resolve.  Weak scaling: 256 spp per GPU (sppTotal = 256*N).

--impl reference: Strelka has no CPU renderer and its OptiX/MDL path cannot be built here,
so the reference arm times oracle/ (the CPU restatement of the same integrator, all host threads) on a
bounded sample of the same workload, on rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, SPP, DEPTH = 1024, 1024, 256, 4
WORKLOAD = "C2 cornell 1024x1024 256spp depth4 (BASELINE.json configs[1])"


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # samples under load = the upper half of the clock readings (the sampler also sees idle gaps)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------
def run_reference(args) -> None:
    """CPU arm: the oracle on the host cores, bounded sample of the same workload, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle
    from strelka_b200.scenes import make_cornell

    cores = os.cpu_count() or 1
    scene, settings, _ = make_cornell(W, H, spp_total=SPP, depth=DEPTH)
    osc = pyoracle.OracleScene(scene)
    sample_spp = 2  # bounded sample per step: 2 of the 256 samples of every pixel (sample indices 0,1)
    times, rays = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, _, _, cnt = osc.render(settings, W, H, sample_spp, threads=cores)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            rays = cnt["radiance_rays"] + cnt["shadow_rays"]
    ms = 1e3 * sum(times) / len(times)
    value = rays / (ms * 1e-3) / 1e6
    sample = f"{sample_spp} of {SPP} spp over all {W}x{H} pixels per step (throughput is spp-independent)"
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "width": W, "height": H, "spp": SPP, "depth": DEPTH, "sample": sample},
        "spp_mpix_per_s": W * H * sample_spp / (ms * 1e-3) / 1e6,
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference OptiX/MDL path unbuildable here (SURVEY 8c); this is the CPU oracle port of the same integrator",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import numpy as np
    import torch

    from strelka_b200 import BufferDesc, BufferFormat, RenderFactory, RenderType, SharedContext
    from strelka_b200.scenes import make_cornell

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 backend has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    spp_total = SPP * world  # weak scaling: 256 samples per pixel per GPU
    scene, settings, _ = make_cornell(W, H, spp_total=spp_total, depth=DEPTH)
    settings.setAs("render/b200/sampleOffset", rank)
    settings.setAs("render/b200/sampleStride", world)

    def make_render(**kw):
        r = RenderFactory.createRender(RenderType.eCompute, device=local, **kw)
        r.setScene(scene)
        r.setSharedContext(SharedContext(mSettingsManager=settings))
        r.init()
        return r

    render = make_render()
    stream = torch.cuda.Stream()  # a dedicated stream shared by the render, the timing events and NCCL
    torch.cuda.set_stream(stream)
    render.set_stream(stream.cuda_stream)
    buf = render.createBuffer(BufferDesc(W, H, BufferFormat.FLOAT4))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        """device-timed step: inputs (scene, BVH, camera, settings) already resident"""
        render.reset_accumulation()
        render.render_iterations(buf, SPP)
        if dist is not None:
            s_acc = render.accum_tensor_nosync()
            dist.all_reduce(s_acc)
            render.resolve(buf, spp_total)

    # first call uploads the scene and builds the BVH (not part of `value`)
    render.render_iterations(buf, 1)
    render.synchronize()
    build_ms = render.counters()["build_ms"]

    for _ in range(args.warmup):
        one_step()
    barrier()
    render.reset_counters()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(1)  # L2 flush between timed iterations (outside the event bracket)
        a.record(stream)
        one_step()
        b.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    c = render.counters()
    rays_step = (c["radiance_rays"] + c["shadow_rays"]) / args.steps
    launches_step = c["kernel_launches"] / args.steps

    # ---- e2e: host buffers in, host image out, through the public API (wall clock) ----------------
    view = scene.view(pinned=True)
    h2d = scene.host_bytes() + 16 * 4 + 4 + 96
    d2h = W * H * 16
    e2e_times = []
    for i in range(1 + max(1, args.steps // 2)):
        barrier()
        t0 = time.perf_counter()
        render.upload_scene_view(view)  # H2D of all scene arrays + on-device BVH build
        render.render_iterations(buf, SPP)
        if dist is not None:
            dist.all_reduce(render.accum_tensor_nosync())
            render.resolve(buf, spp_total)
        img = buf.map()  # blocking D2H into the pinned mirror
        _ = float(img[0, 0, 0])
        barrier()
        if i > 0:
            e2e_times.append(time.perf_counter() - t0)
    e2e_s = sum(e2e_times) / len(e2e_times)

    # ---- max over ranks / sums over ranks -----------------------------------------------------------
    if dist is not None:
        t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(t[0]), float(t[1])
        r_ = torch.tensor([rays_step, launches_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(r_)
        rays_total, launches_total = float(r_[0]), float(r_[1])
    else:
        rays_total, launches_total = rays_step, launches_step

    roofline = cpu_baseline = None
    if rank == 0:
        # ---- roofline of the dominant kernel: one extra, separately instrumented step --------------
        hbm, peak_src = measured_peaks()
        prof = make_render(traversal_stats=True, stage_timers=True)
        pbuf = prof.createBuffer(BufferDesc(W, H, BufferFormat.FLOAT4))
        prof.render_iterations(pbuf, 1)
        prof.reset_accumulation()
        prof.reset_counters()
        prof.render_iterations(pbuf, 32)  # 32 of the 256 iterations: same rays per iteration
        pc = prof.counters()
        # the timed pass of the stage timers must not carry the statistics atomics: time again without them
        timed = make_render(stage_timers=True)
        tbuf = timed.createBuffer(BufferDesc(W, H, BufferFormat.FLOAT4))
        timed.render_iterations(tbuf, 1)
        timed.reset_accumulation()
        timed.reset_counters()
        timed.render_iterations(tbuf, 32)
        tc = timed.counters()
        stages = ("raygen", "extend", "shade", "shadow", "accumulate", "resolve")
        stage_ms = dict(zip(stages, tc["stage_ms"]))
        stage_n = dict(zip(stages, tc["stage_launches"]))
        top = max(stages, key=lambda s: stage_ms[s])
        rr, sr = max(pc["radiance_rays"], 1), max(pc["shadow_rays"], 1)
        per_ray = {
            "extend": 80.0 * pc["nodes_visited"] / rr + 48.0 * pc["tris_tested"] / rr + 64.0 * pc["segs_tested"] / rr + 48.0,
            "shadow": 80.0 * pc["nodes_visited_shadow"] / sr + 48.0 * pc["tris_tested_shadow"] / sr + 64.0 * pc["segs_tested_shadow"] / sr + 44.0,
        }
        trav = top if top in per_ray else "extend"  # the roofline is defined for the traversal kernels
        n_rays = tc["radiance_rays"] if trav == "extend" else tc["shadow_rays"]
        achieved = per_ray[trav] * n_rays / (stage_ms[trav] * 1e-3) / 1e9 if stage_ms[trav] > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("c2", {}).get(f"k_{trav}", {}).get("dram_bytes_per_launch")
        roofline = {
            "bound": "hbm", "kernel": f"k_{trav}" + (" (k_primary for the camera rays + k_extend_simple)" if trav == "extend" else ""), "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
            "traffic": traffic, "traffic_unit": "bytes per launch (ncu --set full, profiles/ncu_traffic.json)",
            "alg_bytes_per_launch": per_ray[trav] * n_rays / max(stage_n[trav], 1),
            "peak_source": peak_src, "alg_bytes_per_ray": per_ray[trav],
            "avg_launch_ms": stage_ms[trav] / max(stage_n[trav], 1), "launches": stage_n[trav],
            "nodes_per_ray": pc["nodes_visited"] / rr, "tris_per_ray": pc["tris_tested"] / rr,
            "stage_ms_share": {s: stage_ms[s] / max(sum(stage_ms.values()), 1e-9) for s in stages},
            "note": "scene fits L1/L2 (36 triangles): traversal is issue-bound, bytes are algorithmic (SURVEY 8d), not DRAM traffic",
        }
        pbuf.destroy(); prof.destroy(); tbuf.destroy(); timed.destroy()
        # ---- CPU baseline: the oracle on the host cores, bounded sample ------------------------------
        from oracle import pyoracle

        cores = os.cpu_count() or 1
        osc = pyoracle.OracleScene(scene)
        sample_spp = 4
        t0 = time.perf_counter()
        _, _, _, cnt = osc.render(settings, W, H, sample_spp, threads=cores)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": (cnt["radiance_rays"] + cnt["shadow_rays"]) / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                        "sample": f"{sample_spp} of {SPP} spp over all {W}x{H} pixels ({dt:.1f} s)"}

    if rank == 0:
        value = rays_total / (dev_ms * 1e-3) / 1e6
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": W, "height": H, "spp_per_gpu": SPP, "spp_total": spp_total, "depth": DEPTH,
                       "parallelism": f"sample-stride x{world} + NCCL all-reduce of S" if world > 1 else "single GPU",
                       "l2": "explicit 256 MiB flush between timed steps; per-batch path state (32 spp x 1 Mpix x 180 B = 5.8 GB) also exceeds L2"},
            "spp_mpix_per_s": W * H * spp_total / (dev_ms * 1e-3) / 1e6,
            "rays_per_step": rays_total, "wall_ms_per_step": 1e3 * t_wall / args.steps, "bvh_build_ms": build_ms,
            "e2e": {"value": rays_total / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s, "includes": "scene upload from pinned host memory, on-device BVH build, render, D2H of the image"},
            "gpu_launches": launches_total * args.steps,
            "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    buf.destroy()
    render.destroy()
    if dist is not None:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
