"""Scratch: how much of the four-matcher step is the target build?  Device-resident 1M/1M match, W matchers in flight,
with the target set on every match (index rebuilt) or only once (index kept)."""
import sys, threading
sys.path.insert(0, ".")
import numpy as np, torch
import libwave_b200 as W
from libwave_b200 import synth
src, tgt, nrm = synth.scan_pair(1_000_000, return_normals=True)
src, tgt, nrm = (synth.to_xyzw(a) for a in (src, tgt, nrm))
n = src.shape[0]
dev = torch.device("cuda:0")
d = [torch.from_numpy(a).to(dev) for a in (src, tgt, nrm)]
for workers in (1, 4):
    streams = [torch.cuda.Stream(device=dev) for _ in range(workers)]
    ms_ = [W.ICPMatcher(W.ICPMatcherParams(res=-1, estimator=W.EST_POINT_TO_PLANE), device=0, stream=s.cuda_stream) for s in streams]
    for mode in ("rebuild", "keep"):
        def work(i, count):
            torch.cuda.set_device(dev)
            m = ms_[i]
            for k in range(count):
                m.setRefDevice(d[0].data_ptr(), n)
                if mode == "rebuild" or k == 0:
                    m.setTargetDevice(d[1].data_ptr(), n); m.setTargetNormalsDevice(d[2].data_ptr(), n)
                assert m.match()
        def round_of(count):
            th = [threading.Thread(target=work, args=(i, count)) for i in range(workers)]
            [t.start() for t in th]; [t.join() for t in th]
        round_of(3)
        res = []
        timer = torch.cuda.Stream(device=dev)
        for _ in range(5):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(timer); round_of(10); e1.record(timer); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / (10 * workers))
        print(f"workers {workers} target {mode}: {np.median(res):.4f} ms per match", flush=True)
