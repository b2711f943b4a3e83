import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import numpy as np
from oracle import pyoracle
from strelka_b200 import BufferDesc, BufferFormat, RenderFactory, RenderType, SharedContext
from strelka_b200.scenes import make_hair
n_strands = int(sys.argv[1]) if len(sys.argv) > 1 else 62500
s, st, (w, h) = make_hair(1024, 1024, 1024, material="hair", n_strands=n_strands)
r = RenderFactory.createRender(RenderType.eCompute); r.setScene(s); r.setSharedContext(SharedContext(mSettingsManager=st)); r.init()
buf = r.createBuffer(BufferDesc(w, h, BufferFormat.FLOAT4))
o = pyoracle.OracleScene(s)
for depth in (1, 2, 3, 6):
  st.setAs("render/pt/depth", depth)
  r.reset_accumulation(); r.render_iterations(buf, 1); img = buf.map().copy()
  print("==== depth", depth)
  x0, y0, ww, wh = 384, 320, 256, 256
  xs, ys = np.meshgrid(np.arange(x0, x0 + ww), np.arange(y0, y0 + wh)); xs, ys = xs.reshape(-1), ys.reshape(-1)
  ref = o.path_radiance(st, w, h, xs, ys, np.zeros(len(xs))).reshape(wh, ww, 3).astype(np.float64)
  got = img[y0:y0 + wh, x0:x0 + ww, :3].astype(np.float64)
  d = np.linalg.norm(got - ref, axis=-1); n = np.linalg.norm(ref, axis=-1)
  print("rel rmse", np.sqrt(((got - ref) ** 2).mean() / (ref ** 2).mean()), "max got", got.max(), "max ref", ref.max())
  rel = d / np.maximum(n, 1e-9); lit = n > 0
  print("lit", lit.sum(), ">1e-3", (rel[lit] > 1e-3).sum(), ">1e-2", (rel[lit] > 1e-2).sum(), ">0.5", (rel[lit] > 0.5).sum(), "got-only", ((d > 0) & ~lit).sum())
  for i in np.argsort(-d.reshape(-1))[:12]:
      print(i // ww, i % ww, got.reshape(-1, 3)[i], ref.reshape(-1, 3)[i])
