import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import libwave_b200 as W
from libwave_b200 import synth
rings, az = synth.SIZES[1_000_000]
scan = synth.velodyne_scan(rings, az, None, synth.SOURCE_SEED, n_points=1_000_000)
big = synth.map_cloud(5, 1_000_000)
m = W.NDTMatcher(W.NDTMatcherParams(res=0.5))
ts = []
for r in range(3):
    t0 = time.perf_counter(); m.setup(scan, big); ok = m.match(); ts.append(time.perf_counter() - t0)
print(json.dumps({"match_ms": 1e3 * min(ts), "iters": m.iterations, "passes": m.stats()["derivative_passes"]}))
