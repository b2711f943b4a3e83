#!/bin/bash
# run the scene report for every library variant under gpurun_variants/ (on the GPU box)
mkdir -p gpurun_out
for lib in gpurun_variants/lib_*.so; do
  n=$(basename $lib .so)
  STRELKA_B200_LIB=$PWD/$lib timeout 300 python tools/scene_report.py "$@" > gpurun_out/ab_$n.log 2>&1
  python - "$n" <<'PY'
import json,sys
n=sys.argv[1]
for l in open(f"gpurun_out/ab_{n}.log"):
    try: d=json.loads(l)
    except Exception: print(n, l.strip()[:200]); continue
    print(f"{n:28s} {d['config']} mrays {d['mrays_s']:8.1f} gen {d['stage_ms']['raygen']:6.2f} ext {d['stage_ms']['extend']:7.2f} shd {d['stage_ms']['shade']:6.2f} sdw {d['stage_ms']['shadow']:7.2f} nodes/ray {d.get('nodes_per_ray')} tris/ray {d.get('tris_per_ray')} frac ext {d.get('extend_roofline_frac')} sdw {d.get('shadow_roofline_frac')}")
PY
done
