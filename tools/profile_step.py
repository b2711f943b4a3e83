"""Small fixed workload for ncu captures: Cornell 1024x1024, a few wavefront batches.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
    ncu --set full --clock-control none --import-source on -k regex:k_shade -s 4 -c 1 -o gpurun_out/shade python tools/profile_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strelka_b200 import BufferDesc, BufferFormat, RenderFactory, RenderType, SharedContext  # noqa: E402
from strelka_b200.scenes import make_cornell  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scene, settings, (w, h) = make_cornell(1024, 1024, spp_total=256)
r = RenderFactory.createRender(RenderType.eCompute)
r.setScene(scene)
r.setSharedContext(SharedContext(mSettingsManager=settings))
r.init()
buf = r.createBuffer(BufferDesc(w, h, BufferFormat.FLOAT4))
r.render_iterations(buf, iters)
r.synchronize()
print(r.counters())
