"""A/B of the single-kernel path tracer against the wavefront form on small scenes (GPU).
Usage: python tools/ab_fused.py [--spp N] ; honours STRELKA_B200_LIB.  Prints device ms and Mrays/s for both forms
and checks that the two images are bit-identical."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strelka_b200 import BufferDesc, BufferFormat, RenderFactory, RenderType, SharedContext  # noqa: E402
from strelka_b200.scenes import make_cornell  # noqa: E402


def run(fused, spp, max_batch=0):
    scene, settings, (w, h) = make_cornell(1024, 1024, 256)
    r = RenderFactory.createRender(RenderType.eCompute, fused_small=fused, max_batch_paths=max_batch)
    r.setScene(scene)
    r.setSharedContext(SharedContext(mSettingsManager=settings))
    r.init()
    buf = r.createBuffer(BufferDesc(w, h, BufferFormat.FLOAT4))
    r.render_iterations(buf, 4)
    r.synchronize()
    best = 1e30
    for _ in range(3):
        r.reset_accumulation()
        r.reset_counters()
        r.render_iterations(buf, spp)
        r.synchronize()
        c = r.counters()
        best = min(best, c["render_ms"])
    img = buf.map().copy()
    buf.unmap()
    rays = c["radiance_rays"] + c["shadow_rays"]
    buf.destroy()
    r.destroy()
    return img, best, rays


if __name__ == "__main__":
    spp = int(sys.argv[sys.argv.index("--spp") + 1]) if "--spp" in sys.argv else 64
    out = {"lib": os.environ.get("STRELKA_B200_LIB", "default"), "spp": spp}
    img_w, ms_w, rays_w = run(False, spp)
    out["wavefront"] = {"ms": round(ms_w, 2), "mrays_s": round(rays_w / ms_w / 1e3, 1)}
    for mb in (0,):
        img_f, ms_f, rays_f = run(True, spp, mb)
        out["fused"] = {"ms": round(ms_f, 2), "mrays_s": round(rays_f / ms_f / 1e3, 1), "rays_equal": rays_f == rays_w,
                        "image_equal": bool(np.array_equal(img_w, img_f))}
    print(json.dumps(out), flush=True)
