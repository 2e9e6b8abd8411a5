"""Scratch driver: one 1M-point point-to-plane match with the tiled kernel, printing the fallback share."""
import sys
import numpy as np
sys.path.insert(0, ".")
import libwave_b200 as W
from libwave_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
src, tgt, nrm = synth.scan_pair(n, return_normals=True)
m = W.ICPMatcher(W.ICPMatcherParams(res=-1, estimator=W.EST_POINT_TO_PLANE))
m.set_profiling(True)
for rep in range(3):
    m.setRef(src); m.setTarget(tgt); m.setTargetNormals(nrm)
    ok = m.match()
st = m.stats()
print(f"ok={ok} iters={m.iterations} pairs={st['pairs']} fallback={st['fallback_queries']} "
      f"({100.0 * st['fallback_queries'] / max(1, st['pairs']):.2f} %) corr/iter={st['iterate_ms'] / max(1, st['iterate_launches']) * 1e3:.1f} us "
      f"build={st['build_ms']:.3f} ms")
