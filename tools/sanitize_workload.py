"""Small renders of every code path (wavefront, fused small-scene kernel, instrumented traversal, curves) for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_workload.py
    compute-sanitizer --tool racecheck python tools/sanitize_workload.py
Round 1: 0 errors / 0 hazards on B200.  Round 2 adds the persistent triangle kernels (parking queue, 256-bit node
loads, shared-memory octant table), textures and the hair BSDF."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from strelka_b200 import BufferDesc, BufferFormat, RenderFactory, RenderType, SharedContext
from strelka_b200.scenes import make_cornell, make_hair, make_kitchen
from util import random_scene
def run(scene, settings, w, h, n, **kw):
    r = RenderFactory.createRender(RenderType.eCompute, **kw)
    r.setScene(scene); r.setSharedContext(SharedContext(mSettingsManager=settings)); r.init()
    buf = r.createBuffer(BufferDesc(w, h, BufferFormat.FLOAT4))
    r.render_iterations(buf, n); img = buf.map().copy(); buf.destroy(); r.destroy()
    return float(img[..., :3].mean())
s, st, _ = make_cornell(96, 64, 6); print("cornell", run(s, st, 96, 64, 6))
print("cornell fused", run(s, st, 96, 64, 6, fused_small=True))
s, st = random_scene(seed=2); st.setAs("render/pt/sppTotal", 3); st.setAs("render/pt/depth", 6); print("random", run(s, st, 70, 50, 3, traversal_stats=True))
s, st, (w, h) = make_hair(48, 48, 2, n_strands=300, segments=8); print("hair", run(s, st, w, h, 2))
s, st, (w, h) = make_hair(48, 48, 2, n_strands=300, segments=8, material="hair"); print("hair bsdf", run(s, st, w, h, 2))
s, st, _ = make_kitchen(64, 36, 2, n_props=24, subdiv=2); print("kitchen (persistent triangle kernels)", run(s, st, 64, 36, 2))
s, st, _ = make_kitchen(64, 36, 2, n_props=24, subdiv=2, textured=True); print("kitchen textured", run(s, st, 64, 36, 2, traversal_stats=True))
