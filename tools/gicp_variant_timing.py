"""Scratch: GICP 500k timing per tuning build (one process per WAVECU_SO variant)."""
import json, os, subprocess, sys, time
sys.path.insert(0, ".")
import numpy as np
CACHE = "/tmp/gicp_clouds.npz"
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    import libwave_b200 as W
    from libwave_b200 import synth
    d = np.load(CACHE)
    s4, t4 = synth.to_xyzw(d["src"]), synth.to_xyzw(d["tgt"])
    m = W.GICPMatcher(W.GICPMatcherParams(res=-1))
    m.set_profiling(True)
    best = None
    for r in range(3):
        m.setRef(s4); m.setTarget(t4)
        t0 = time.perf_counter(); ok = m.match(); t1 = time.perf_counter()
        st = m.stats()
        cur = {"match_ms": round(1e3 * (t1 - t0), 2), "iters": m.iterations, "evals": st["evaluations"],
               "cost_us": round(1e3 * st["cost_kernel_ms"] / max(1, st["cost_kernel_launches"]), 2),
               "T03": float(m.getResult()[0, 3])}
        if best is None or cur["cost_us"] < best["cost_us"]:
            best = cur
    print(os.environ.get("WAVECU_SO", "libwavecu.so"), json.dumps(best))
else:
    if not os.path.exists(CACHE):
        from libwave_b200 import synth
        src, tgt = synth.scan_pair(500_000)
        np.savez(CACHE, src=src, tgt=tgt)
    for so in (sys.argv[1:] or ["libwavecu.so"]):
        subprocess.run([sys.executable, __file__, "--one"], env=dict(os.environ, WAVECU_SO=so))
