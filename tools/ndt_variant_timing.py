"""Scratch: NDT 1M-vs-5M timing per tuning build: clouds generated once and cached in /tmp, then one
process per WAVECU_SO variant.   python tools/ndt_variant_timing.py [variant.so ...]"""
import json, os, subprocess, sys, time
sys.path.insert(0, ".")
import numpy as np

CACHE = "/tmp/ndt_clouds.npz"
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    import libwave_b200 as W
    from libwave_b200 import synth
    d = np.load(CACHE)
    scan4, big4 = synth.to_xyzw(d["scan"]), synth.to_xyzw(d["big"])
    m = W.NDTMatcher(W.NDTMatcherParams(res=0.5))
    m.set_profiling(True)
    best = None
    for r in range(4):
        t0 = time.perf_counter(); m.setup(scan4, big4); t1 = time.perf_counter(); ok = m.match(); t2 = time.perf_counter()
        st = m.stats()
        cur = {"match_ms": round(1e3 * (t2 - t1), 2), "iters": m.iterations, "passes": st["derivative_passes"],
               "pass_us": round(1e3 * st["derivative_kernel_ms"] / max(1, st["derivative_passes"]), 1)}
        if best is None or cur["pass_us"] < best["pass_us"]:
            best = cur
    print(os.environ.get("WAVECU_SO", "libwavecu.so"), json.dumps(best))
else:
    if not os.path.exists(CACHE):
        from bench import ndt_clouds
        scan, big = ndt_clouds(0)
        np.savez(CACHE, scan=scan, big=big)
    for so in (sys.argv[1:] or ["libwavecu.so"]):
        env = dict(os.environ, WAVECU_SO=so)
        subprocess.run([sys.executable, __file__, "--one"], env=env)
