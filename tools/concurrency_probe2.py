"""Scratch: throughput of W concurrent GICP / NDT matchers on one GPU (BASELINE configs 3 and 4), device-resident."""
import sys, threading
sys.path.insert(0, ".")
import numpy as np, torch
import libwave_b200 as W
from libwave_b200 import synth
dev = torch.device("cuda:0")

def probe(name, make, inputs, reps):
    d = [torch.from_numpy(a).to(dev) for a in inputs]
    n = [a.shape[0] for a in inputs]
    for workers in (1, 2, 4):
        streams = [torch.cuda.Stream(device=dev) for _ in range(workers)]
        ms_ = [make(s.cuda_stream) for s in streams]
        def work(i, count):
            torch.cuda.set_device(dev)
            m = ms_[i]
            for _ in range(count):
                m.setRefDevice(d[0].data_ptr(), n[0]); m.setTargetDevice(d[1].data_ptr(), n[1])
                assert m.match()
        def round_of(count):
            th = [threading.Thread(target=work, args=(i, count)) for i in range(workers)]
            [t.start() for t in th]; [t.join() for t in th]
        round_of(2)
        res = []
        timer = torch.cuda.Stream(device=dev)
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(timer); round_of(reps); e1.record(timer); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / (reps * workers))
        print(f"{name} workers {workers}: {np.median(res):.3f} ms per match", flush=True)
        del ms_

src, tgt = synth.scan_pair(500_000)
probe("gicp500k", lambda st: W.GICPMatcher(W.GICPMatcherParams(res=-1), device=0, stream=st), (synth.to_xyzw(src), synth.to_xyzw(tgt)), 5)
rings, az = synth.SIZES[1_000_000]
scan = synth.velodyne_scan(rings, az, None, synth.SOURCE_SEED, n_points=1_000_000)
big = synth.map_cloud(5, 1_000_000)
probe("ndt1m5m", lambda st: W.NDTMatcher(W.NDTMatcherParams(res=0.5), device=0, stream=st), (synth.to_xyzw(scan), synth.to_xyzw(big)), 5)
