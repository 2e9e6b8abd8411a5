"""Scratch driver: one NDT match of the 1M-vs-5M configuration (for ncu captures)."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import libwave_b200 as W
CACHE = "/tmp/ndt_clouds.npz"
if os.path.exists(CACHE):
    d = np.load(CACHE)
    scan, big = d["scan"], d["big"]
else:
    from bench import ndt_clouds
    scan, big = ndt_clouds(0)
m = W.NDTMatcher(W.NDTMatcherParams(res=0.5))
m.setup(scan, big)
print(m.match(), m.iterations, m.stats())
