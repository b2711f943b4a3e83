"""CPU-oracle throughput on the BASELINE configs (reduced spp, full scene size): the CPU column of BASELINE.md.
Usage: python tools/cpu_report.py [c2 c3 c4 c5] [--spp N]   (test infrastructure: runs the oracle, not the product)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle  # noqa: E402
from strelka_b200.scenes import make_cornell, make_hair, make_instanced, make_kitchen  # noqa: E402

CONFIGS = {
    "c2": lambda: make_cornell(1024, 1024, 256),
    "c3": lambda: make_kitchen(1920, 1080, 2048),
    "c4": lambda: make_hair(1024, 1024, 1024, depth=6),
    "c5": lambda: make_instanced(3840, 2160, 4096),
}

if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    spp = 1
    if "--spp" in sys.argv:
        spp = int(sys.argv[sys.argv.index("--spp") + 1])
        args = [a for a in args if a != str(spp)]
    cores = os.cpu_count() or 1
    for name in args or ["c2", "c3", "c4", "c5"]:
        scene, settings, (w, h) = CONFIGS[name]()
        t0 = time.perf_counter()
        osc = pyoracle.OracleScene(scene)
        build_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        _, _, _, cnt = osc.render(settings, w, h, spp, threads=cores)
        dt = time.perf_counter() - t0
        rays = cnt["radiance_rays"] + cnt["shadow_rays"]
        print(json.dumps({"config": name, "cores": cores, "spp": spp, "bvh2_build_s": round(build_s, 1), "render_s": round(dt, 2),
                          "mrays_s": round(rays / dt / 1e6, 2), "spp_mpix_s": round(w * h * spp / dt / 1e6, 3)}), flush=True)
        osc.close()
