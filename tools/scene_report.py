"""Per-config report on the GPU: build time, Mrays/s, spp*Mpix/s, traversal statistics and roofline fraction
for the BASELINE configs at full scene size (reduced spp).  Usage: python tools/scene_report.py [c2 c3 c4 c5] [--spp N]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strelka_b200 import BufferDesc, BufferFormat, RenderFactory, RenderType, SharedContext  # noqa: E402
from strelka_b200.scenes import make_cornell, make_hair, make_instanced, make_kitchen  # noqa: E402

HBM = 6452.8
if os.path.exists("MEASURED_PEAKS.json"):
    HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]

CONFIGS = {
    "c2": lambda: make_cornell(1024, 1024, 256),
    "c3": lambda: make_kitchen(1920, 1080, 2048),
    "c4": lambda: make_hair(1024, 1024, 1024, depth=6),
    "c5": lambda: make_instanced(3840, 2160, 4096),
}


def run(name, spp):
    t0 = time.time()
    scene, settings, (w, h) = CONFIGS[name]()
    gen_s = time.time() - t0
    out = {"config": name, "width": w, "height": h, "spp_rendered": spp, "gen_s": round(gen_s, 2), "host_mb": round(scene.host_bytes() / 1e6, 1)}
    for mode in ("timed", "stats"):
        r = RenderFactory.createRender(RenderType.eCompute, traversal_stats=(mode == "stats"), stage_timers=True,
                                       curve_split=int(os.environ.get("STRELKA_CURVE_SPLIT", "0")))
        r.setScene(scene)
        r.setSharedContext(SharedContext(mSettingsManager=settings))
        r.init()
        buf = r.createBuffer(BufferDesc(w, h, BufferFormat.FLOAT4))
        r.render_iterations(buf, 1)
        r.synchronize()
        c0 = r.counters()
        r.reset_accumulation()
        r.reset_counters()
        t0 = time.time()
        r.render_iterations(buf, spp)
        r.synchronize()
        wall = time.time() - t0
        c = r.counters()
        if mode == "timed":
            rays = c["radiance_rays"] + c["shadow_rays"]
            wall = max(sum(c["stage_ms"]) * 1e-3, 1e-9)  # device time of the kernels (the host wall clock also sees queue growth)
            out.update(build_ms=round(c0["build_ms"], 1), tris=c["num_triangles"], segs=c["num_segments"], nodes_tri=c["bvh_nodes_tri"],
                       nodes_seg=c["bvh_nodes_curve"], mrays_s=round(rays / wall / 1e6, 1), spp_mpix_s=round(w * h * spp / wall / 1e6, 2),
                       wall_s=round(wall, 3), rays_per_path=round(rays / c["paths"], 2),
                       stage_ms={k: round(v, 2) for k, v in zip(("raygen", "extend", "shade", "shadow", "accumulate", "resolve"), c["stage_ms"])})
            timed = c
        else:
            rr, sr = max(c["radiance_rays"], 1), max(c["shadow_rays"], 1)
            b_ext = 80 * c["nodes_visited"] / rr + 48 * c["tris_tested"] / rr + 64 * c["segs_tested"] / rr + 48
            b_sh = 80 * c["nodes_visited_shadow"] / sr + 48 * c["tris_tested_shadow"] / sr + 64 * c["segs_tested_shadow"] / sr + 44
            ext_ms, sh_ms = timed["stage_ms"][1], timed["stage_ms"][3]
            out.update(nodes_per_ray=round(c["nodes_visited"] / rr, 2), tris_per_ray=round(c["tris_tested"] / rr, 2),
                       segs_per_ray=round(c["segs_tested"] / rr, 2), nodes_per_shadow_ray=round(c["nodes_visited_shadow"] / sr, 2),
                       alg_bytes_per_ray=round(b_ext, 1), stack_overflows=c["stack_overflows"],
                       extend_gbs=round(b_ext * timed["radiance_rays"] / (ext_ms * 1e-3) / 1e9, 1) if ext_ms else None,
                       shadow_gbs=round(b_sh * timed["shadow_rays"] / (sh_ms * 1e-3) / 1e9, 1) if sh_ms else None)
            out["extend_roofline_frac"] = round(out["extend_gbs"] / HBM, 3) if out["extend_gbs"] else None
            out["shadow_roofline_frac"] = round(out["shadow_gbs"] / HBM, 3) if out["shadow_gbs"] else None
        buf.destroy()
        r.destroy()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    spp = 8
    if "--spp" in sys.argv:
        spp = int(sys.argv[sys.argv.index("--spp") + 1])
        args = [a for a in args if a != str(spp)]
    for n in args or ["c2", "c3", "c4", "c5"]:
        run(n, spp)
