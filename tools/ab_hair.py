"""stage times of the C4 step with the hair fibre BSDF for every library variant (tools/ab.sh analogue)"""
import glob, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import bench
    hx = bench.Harness()
    rec = bench.bench_scene(hx, "c4", 2, 3, False, False)
    sh = rec["stage_ms_share"]
    print(f"{sys.argv[1]:24s} mrays {rec['value']:7.1f} ms {rec['ms_per_step']:7.1f} extend {sh['extend']*100:5.1f}% shade {sh['shade']*100:5.1f}% shadow {sh['shadow']*100:5.1f}%")
else:
    for lib in sorted(glob.glob("gpurun_variants/lib_*.so")):
        env = dict(os.environ, STRELKA_B200_LIB=os.path.abspath(lib))
        subprocess.run([sys.executable, __file__, os.path.basename(lib)], env=env)
