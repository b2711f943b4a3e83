import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import numpy as np
import bsdf_pins_common as P
from test_bsdf_pins import run as run_o
from strelka_b200 import _abi, RenderFactory, RenderType
r = RenderFactory.createRender(RenderType.eCompute); r.init()
run_g = lambda m, packed: r.test_bsdf(m, packed)
kw = P.HAIR_CASES[0]
m = P.material(_abi.SB_MATERIAL_HAIR, **kw)
rng = np.random.default_rng(0); n = 200000
k1 = P.unit(rng.normal(size=(n, 3))); k2 = P.unit(rng.normal(size=(n, 3)))
packed = P.pack(P.N_HAIR, P.N_HAIR, P.T_HAIR, k1, rng.random((n, 4)), k2)
se, ee = run_g(m, packed); so, eo = run_o(m, packed)
g_e = ee[:, 3:6].astype(np.float64); g_o = eo[:, 3:6].astype(np.float64)
rel = np.abs(g_e - g_o).max(1) / np.maximum(np.abs(g_o).max(1), 1e-12)
print("eval rel err: median %.2e p99 %.2e p999 %.2e max %.2e" % (np.median(rel), np.percentile(rel, 99), np.percentile(rel, 99.9), rel.max()))
for j in np.argsort(-rel)[:5]:
    print(rel[j], k1[j], k2[j], g_e[j], g_o[j])
pe = np.abs(ee[:, 6] - eo[:, 6]) / np.maximum(eo[:, 6], 1e-12)
print("pdf rel err: median %.2e p99 %.2e max %.2e" % (np.median(pe), np.percentile(pe, 99), pe.max()))
ok = (se[:, 7] == so[:, 7])
w_e = se[ok, 3:6].astype(np.float64); w_o = so[ok, 3:6].astype(np.float64)
relw = np.abs(w_e - w_o).max(1) / np.maximum(np.abs(w_o).max(1), 1e-12)
print("sample weight rel err: median %.2e p99 %.2e max %.2e; event mismatch %d" % (np.median(relw), np.percentile(relw, 99), relw.max(), (~ok).sum()))
dk = np.linalg.norm(se[ok, :3] - so[ok, :3], axis=1)
print("k2 diff: median %.2e p99 %.2e p999 %.2e max %.2e" % (np.median(dk), np.percentile(dk, 99), np.percentile(dk, 99.9), dk.max()))
for j in np.argsort(-dk)[:5]:
    print(dk[j], packed[ok][j, 9:16], se[ok][j, :3], so[ok][j, :3])
