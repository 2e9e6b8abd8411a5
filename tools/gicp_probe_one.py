"""Scratch driver: one GICP match of the 500k configuration (for ncu captures)."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
import libwave_b200 as W
from libwave_b200 import synth
CACHE = "/tmp/gicp_clouds.npz"
if os.path.exists(CACHE):
    d = np.load(CACHE)
    src, tgt = d["src"], d["tgt"]
else:
    src, tgt = synth.scan_pair(500_000)
m = W.GICPMatcher(W.GICPMatcherParams(res=-1))
m.setup(src, tgt)
print(m.match(), m.iterations, m.stats())
