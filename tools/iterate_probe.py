"""Scratch driver for ncu: three device-resident 1M/1M point-to-plane matches on one matcher (the bench workload)."""
import sys
sys.path.insert(0, ".")
import torch
import libwave_b200 as W
from libwave_b200 import synth
src, tgt, nrm = synth.scan_pair(1_000_000, return_normals=True)
d = [torch.from_numpy(synth.to_xyzw(a)).cuda() for a in (src, tgt, nrm)]
m = W.ICPMatcher(W.ICPMatcherParams(res=-1, estimator=W.EST_POINT_TO_PLANE))
n = src.shape[0]
for _ in range(3):
    m.setRefDevice(d[0].data_ptr(), n); m.setTargetDevice(d[1].data_ptr(), n); m.setTargetNormalsDevice(d[2].data_ptr(), n)
    print(m.match(), m.iterations)
