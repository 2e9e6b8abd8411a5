"""Scratch driver: the reference's default path (res = 0.1, multiscale_steps = 3) on the 1M / 200k synthetic pairs."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import libwave_b200 as W
from libwave_b200 import synth
for n in (200_000, 1_000_000):
    src, tgt = synth.scan_pair(n)
    s4, t4 = synth.to_xyzw(src), synth.to_xyzw(tgt)
    m = W.ICPMatcher(W.ICPMatcherParams())   # defaults: res 0.1, multiscale_steps 3
    for rep in range(3):
        t0 = time.perf_counter(); m.setup(s4, t4); ok = m.match(); dt = time.perf_counter() - t0
    print(f"n={n}: default multiscale match ok={ok} iters={m.iterations} {dt*1e3:.2f} ms  T[:3,3]={m.getResult()[:3,3]}")
