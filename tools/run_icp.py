"""Scratch driver: per-configuration ICP timing breakdown on the GPU (not part of the product)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import libwave_b200 as W
from libwave_b200 import synth

for n in (200_000, 1_000_000):
    src, tgt, nrm = synth.scan_pair(n, return_normals=True)
    for est, name in ((W.EST_SVD, "p2p"), (W.EST_POINT_TO_PLANE, "p2plane")):
        m = W.ICPMatcher(W.ICPMatcherParams(res=-1, estimator=est))
        m.set_profiling(True)
        m.setup(src, tgt)
        if est == W.EST_POINT_TO_PLANE:
            m.setTargetNormals(nrm)
        for rep in range(3):
            m.setRef(src); m.setTarget(tgt)
            if est == W.EST_POINT_TO_PLANE:
                m.setTargetNormals(nrm)
            t0 = time.perf_counter(); ok = m.match(); dt = time.perf_counter() - t0
        st = m.stats()
        it = max(1, st["iterate_launches"])
        print(f"n={n} {name}: ok={ok} iters={m.iterations} wall={dt*1e3:.2f}ms total={st['total_ms']:.2f} build={st['build_ms']:.3f} "
              f"corr/iter={st['iterate_ms']/it*1e3:.1f}us red+solve/iter={st['solve_ms']/it*1e3:.1f}us launches={st['kernel_launches']}")
