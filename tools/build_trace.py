"""Phase times of sb_set_scene (upload, flatten, sort, PLOC, BVH8 cut, collapse) on the GPU, cold and warm.
Usage: STRELKA_B200_BUILD_TRACE=1 python tools/build_trace.py [c3 c4 c5]"""
import os
import sys
import time

os.environ.setdefault("STRELKA_B200_BUILD_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_scene  # noqa: E402
from strelka_b200 import RenderFactory, RenderType, SharedContext  # noqa: E402

for key in sys.argv[1:] or ["c3", "c4", "c5"]:
    scene, settings, _ = make_scene(key)
    r = RenderFactory.createRender(RenderType.eCompute)
    r.setScene(scene)
    r.setSharedContext(SharedContext(mSettingsManager=settings))
    r.init()
    view = scene.view(pinned=True)
    for i in range(3):
        print(f"==== {key} sb_set_scene call {i} (pinned host arrays)", file=sys.stderr, flush=True)
        t0 = time.perf_counter()
        r.upload_scene_view(view)
        r.synchronize()
        print(f"==== {key} call {i}: wall {1e3 * (time.perf_counter() - t0):.1f} ms, build_ms {r.counters()['build_ms']:.1f}", file=sys.stderr, flush=True)
    r.destroy()
