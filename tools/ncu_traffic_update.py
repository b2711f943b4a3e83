"""Fold `ncu --set full` raw pages (tools/capture_traversal.sh) into profiles/: one key-metric summary per capture
(profiles/<tag>_<cfg>_<kernel>_full.txt), per-source-line hotspots where the profiled library is at hand, and the
per-launch figures bench.py reads (profiles/ncu_traffic.json: DRAM bytes, DRAM / L2 GB/s, issue-active, lanes).

    python tools/ncu_traffic_update.py r02 c3 c4 c5 [--lib strelka_b200/libstrelka_b200.so]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ncu_summary import KEYS  # noqa: E402

EXTRA = ["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sectors.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
         "l1tex__throughput.avg.pct_of_peak_sustained_active", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
         "dram__bytes.sum.per_second", "sm__cycles_elapsed.avg.per_second"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    lib = sys.argv[sys.argv.index("--lib") + 1] if "--lib" in sys.argv else os.path.join(ROOT, "strelka_b200", "libstrelka_b200.so")
    args = [a for a in args if a != lib]
    tag, cfgs = args[0], args[1:]
    hbm = 6452.8
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        hbm = json.load(open(p))["hbm_gbs"]
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    traffic["_doc"] = ("per-launch figures of the shipped kernels from `ncu --set full --clock-control none` captures (tools/capture_traversal.sh, "
                       "tools/ncu_traffic_update.py): dram__bytes_read.sum + dram__bytes_write.sum, DRAM and L2 (lts__t_sectors x 32 B) throughput, "
                       "smsp__issue_active, active lanes per instruction; read by bench.py for roofline.traffic and the ncu_* keys")
    for cfg in cfgs:
        for k in ("k_extend", "k_shadow", "k_shade"):
            raw = os.path.join(ROOT, "gpurun_out", f"{tag}_{cfg}_{k}_raw.csv")
            if not os.path.exists(raw):
                continue
            rows = list(csv.reader(open(raw)))
            if len(rows) < 3:
                continue
            hdr, units, vals = rows[0], rows[1], rows[2]
            get = lambda name: (vals[hdr.index(name)], units[hdr.index(name)]) if name in hdr else (None, None)  # noqa: E731

            def num(name):
                v, u = get(name)
                return None if v in (None, "") else float(v.replace(",", "")) * UNIT.get(u, 1.0)

            name = vals[hdr.index("Kernel Name")]
            out = os.path.join(ROOT, "profiles", f"{tag}_{cfg}_{k}_full.txt")
            with open(out, "w") as f:
                f.write(f"# ncu --set full --clock-control none, one launch: {name}\n# workload: tools/profile_step.py {cfg} 4 "
                        f"(first launch of the kernel = bounce 1 of a 4-sample batch at the config's full size)\n")
                for key in KEYS + EXTRA:
                    v, u = get(key)
                    if v is not None:
                        f.write(f"{key:90s} {v} {u}\n")
            dur = num("gpu__time_duration.sum")
            dram = (num("dram__bytes_read.sum") or 0.0) + (num("dram__bytes_write.sum") or 0.0)
            l2 = (num("lts__t_sectors.sum") or 0.0) * 32.0
            rec = {"dram_bytes_per_launch": dram, "launch_ms": dur * 1e3, "dram_gbs": dram / dur / 1e9, "dram_frac_of_peak": dram / dur / 1e9 / hbm,
                   "l2_gbs": l2 / dur / 1e9, "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "warp_lanes_active": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
                   "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num("lts__t_sector_hit_rate.pct"),
                   "l1_data_pipe_pct": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                   "alu_pipe_pct": num("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                   "launch": f"{name.split('(')[0].replace('void ', '')}, bounce 1 of a 4-sample batch", "source": f"profiles/{tag}_{cfg}_{k}_full.txt"}
            traffic.setdefault(cfg, {})[k] = rec
            src = os.path.join(ROOT, "gpurun_out", f"{tag}_{cfg}_{k}_source.csv")
            if os.path.exists(src):
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hotspots.py"), src, lib, "--top", "30"], capture_output=True, text=True)
                if r.returncode == 0 and "not the profiled build" not in r.stdout:
                    open(os.path.join(ROOT, "profiles", f"{tag}_{cfg}_{k}_hotspots.txt"), "w").write(r.stdout)
    json.dump(traffic, open(tpath, "w"), indent=1)
    print(json.dumps({c: {k: {m: (round(v, 3) if isinstance(v, float) else v) for m, v in r.items() if m not in ("launch", "source")} for k, r in t.items()}
                      for c, t in traffic.items() if c != "_doc"}, indent=1))


if __name__ == "__main__":
    main()
