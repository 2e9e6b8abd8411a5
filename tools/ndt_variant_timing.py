"""Scratch: NDT 1M-vs-5M timing split into upload (setup) and match()."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import libwave_b200 as W
from libwave_b200 import synth
rings, az = synth.SIZES[1_000_000]
scan = synth.velodyne_scan(rings, az, None, synth.SOURCE_SEED, n_points=1_000_000)
big = synth.map_cloud(5, 1_000_000)
scan4, big4 = synth.to_xyzw(scan), synth.to_xyzw(big)
m = W.NDTMatcher(W.NDTMatcherParams(res=0.5))
best = None
for r in range(3):
    t0 = time.perf_counter(); m.setup(scan4, big4); t1 = time.perf_counter(); ok = m.match(); t2 = time.perf_counter()
    cur = {"setup_ms": 1e3 * (t1 - t0), "match_ms": 1e3 * (t2 - t1), "iters": m.iterations, "passes": m.stats()["derivative_passes"]}
    if best is None or cur["match_ms"] < best["match_ms"]:
        best = cur
print(json.dumps(best))
