"""Scratch: do kernels slow down while an H2D copy is in flight?  Device-resident match with and without a concurrent
pinned-host -> device copy on another stream (WAVECU_TIMELINE prints the phase times)."""
import os, sys, time
os.environ["WAVECU_TIMELINE"] = "1"; os.environ["WAVECU_NO_GRAPH"] = "1"
sys.path.insert(0, ".")
import numpy as np, torch
import libwave_b200 as W
from libwave_b200 import synth
src, tgt, nrm = synth.scan_pair(1_000_000, return_normals=True)
src, tgt, nrm = (synth.to_xyzw(a) for a in (src, tgt, nrm))
d = [torch.from_numpy(a).cuda() for a in (src, tgt, nrm)]
big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
dbig = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
side = torch.cuda.Stream()
m = W.ICPMatcher(W.ICPMatcherParams(res=-1, estimator=W.EST_POINT_TO_PLANE))
m.set_profiling(True)
n = src.shape[0]
for mode in ("quiet", "h2d", "quiet", "h2d", "d2h"):
    torch.cuda.synchronize()
    if mode == "h2d":
        with torch.cuda.stream(side):
            dbig.copy_(big, non_blocking=True)
    if mode == "d2h":
        with torch.cuda.stream(side):
            big.copy_(dbig, non_blocking=True)
    time.sleep(0.0005)
    print(mode, file=sys.stderr)
    m.setRefDevice(d[0].data_ptr(), n); m.setTargetDevice(d[1].data_ptr(), n); m.setTargetNormalsDevice(d[2].data_ptr(), n)
    m.match()
    torch.cuda.synchronize()
    st = m.stats()
    print(f"   build {st['build_ms']:.3f} iterate {st['iterate_ms']:.3f} ({st['iterate_launches']} launches) total {st['total_ms']:.3f} ms", file=sys.stderr)
# a plain streaming kernel (device-to-device copy of 256 MB) with and without the H2D copy
a = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
for mode in ("quiet", "h2d", "quiet", "h2d"):
    torch.cuda.synchronize()
    if mode == "h2d":
        with torch.cuda.stream(side):
            dbig.copy_(big, non_blocking=True)
    time.sleep(0.0005)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    print(f"d2d 256 MB {mode}: {e0.elapsed_time(e1):.3f} ms", file=sys.stderr)
