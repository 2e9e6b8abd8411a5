"""Scratch: stream timeline of one match (WAVECU_TIMELINE=1), host-resident vs device-resident inputs."""
import os, sys, time
os.environ["WAVECU_TIMELINE"] = "1"
sys.path.insert(0, ".")
import numpy as np, torch
import libwave_b200 as W
from libwave_b200 import synth
src, tgt, nrm = synth.scan_pair(1_000_000, return_normals=True)
src, tgt, nrm = (synth.to_xyzw(a) for a in (src, tgt, nrm))
h = [torch.from_numpy(a).pin_memory() for a in (src, tgt, nrm)]
d = [torch.from_numpy(a).cuda() for a in (src, tgt, nrm)]
m = W.ICPMatcher(W.ICPMatcherParams(res=-1, estimator=W.EST_POINT_TO_PLANE))
m.set_profiling(True)
n = src.shape[0]
for rep in range(3):   # target first: its index build overlaps the source upload
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m.setTarget(h[1].numpy()); m.setRef(h[0].numpy()); m.setTargetNormals(h[2].numpy()); t3 = time.perf_counter()
    ok = m.match(); t4 = time.perf_counter()
    print(f"host, target first: set* {1e3*(t3-t0):.3f} match {1e3*(t4-t3):.3f} total {1e3*(t4-t0):.3f} ms", file=sys.stderr)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m.setRef(h[0].numpy()); t1 = time.perf_counter(); m.setTarget(h[1].numpy()); t2 = time.perf_counter(); m.setTargetNormals(h[2].numpy()); t3 = time.perf_counter()
    ok = m.match(); t4 = time.perf_counter()
    print(f"host: setRef {1e3*(t1-t0):.3f} setTarget {1e3*(t2-t1):.3f} setNormals {1e3*(t3-t2):.3f} match {1e3*(t4-t3):.3f} total {1e3*(t4-t0):.3f} ms", file=sys.stderr)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m.setRefDevice(d[0].data_ptr(), n); m.setTargetDevice(d[1].data_ptr(), n); m.setTargetNormalsDevice(d[2].data_ptr(), n); t3 = time.perf_counter()
    ok = m.match(); t4 = time.perf_counter()
    print(f"device: set* {1e3*(t3-t0):.3f} match {1e3*(t4-t3):.3f} total {1e3*(t4-t0):.3f} ms", file=sys.stderr)
