"""Scratch driver for compute-sanitizer: small instances of every round-2 kernel path."""
import sys
sys.path.insert(0, ".")
import numpy as np
import libwave_b200 as W
from libwave_b200 import batch, synth

src, tgt, nrm = synth.scan_pair(10_000, return_normals=True)
for mode in (W.SEARCH_TREE, W.SEARCH_TILED):
    for est in (W.EST_SVD, W.EST_POINT_TO_PLANE):
        m = W.ICPMatcher(W.ICPMatcherParams(res=-1, estimator=est))
        m.set_search(mode)
        m.setup(src, tgt)
        if est == W.EST_POINT_TO_PLANE:
            m.setTargetNormals(nrm)
        ok = m.match()
        m.estimateInfo()
        print("icp", mode, est, ok, m.iterations, m.stats()["fallback_queries"])
m = W.ICPMatcher(W.ICPMatcherParams(res=0.2, multiscale_steps=2))
m.set_search(W.SEARCH_TILED)
m.setup(src, tgt)
print("multiscale tiled", m.match(), m.iterations)
sb = batch.ScanBatch(W.ICPMatcherParams(res=-1), devices=[0], workers_per_device=2)
sources, target = synth.scan_batch(10_000, 0, ids=[0, 1, 2])
sb.set_map(target)
recs = sb.match([synth.to_xyzw(s) for s in sources], with_info=True)
print("batch", [r.iterations for r in recs])
g = W.GICPMatcher(W.GICPMatcherParams(res=-1))
g.setup(src[:4000], tgt[:4000])
print("gicp", g.match(), g.iterations)
for res in (0.5, 0.02):   # dense table / hash table
    n = W.NDTMatcher(W.NDTMatcherParams(res=res))
    n.setup(src, tgt)
    print("ndt", res, n.match(), n.iterations, n.stats()["n_cells"])

# four matchers in flight on one device, one host thread each (bench.py's structure), host and device inputs
import threading
import torch
xs, xt, xn = (synth.to_xyzw(a) for a in (src, tgt, nrm))
hp = [tuple(torch.from_numpy(a).pin_memory() for a in (xs, xt, xn)) for _ in range(4)]
crew = [W.ICPMatcher(W.ICPMatcherParams(res=-1, estimator=W.EST_POINT_TO_PLANE)) for _ in range(4)]
out = []


def work(i):
    mm = crew[i]
    for _ in range(3):
        mm.setRef(hp[i][0].numpy()); mm.setTarget(hp[i][1].numpy()); mm.setTargetNormals(hp[i][2].numpy())
        out.append((i, mm.match(), mm.iterations))


th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
[t.start() for t in th]; [t.join() for t in th]
print("crew", sorted(out))
