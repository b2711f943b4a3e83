"""Small fixed workload for ncu captures.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py c2 8
    ncu --set full --clock-control none --import-source on -k regex:k_shade -s 4 -c 1 -o gpurun_out/shade python tools/profile_step.py c3 2
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strelka_b200 import BufferDesc, BufferFormat, RenderFactory, RenderType, SharedContext  # noqa: E402
from strelka_b200.scenes import make_cornell, make_hair, make_instanced, make_kitchen  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
make = {"c2": lambda: make_cornell(1024, 1024, 256), "c3": lambda: make_kitchen(1920, 1080, 2048),
        "c4": lambda: make_hair(1024, 1024, 1024, depth=6), "c5": lambda: make_instanced(3840, 2160, 4096)}[cfg]
scene, settings, (w, h) = make()
r = RenderFactory.createRender(RenderType.eCompute)
r.setScene(scene)
r.setSharedContext(SharedContext(mSettingsManager=settings))
r.init()
buf = r.createBuffer(BufferDesc(w, h, BufferFormat.FLOAT4))
r.render_iterations(buf, iters)
r.synchronize()
print(r.counters())
