"""Scratch: the reference-facing C++ MultiMatcher (tools/native/multimatcher_bench, no Python in the measured process)
against the same jobs driven from Python threads: 1M/1M and 200k/200k SVD ICP + estimateInfo per job, unpinned host
clouds, 1 / 2 / 4 workers on one GPU."""
import os, subprocess, sys, threading, time
sys.path.insert(0, ".")
import numpy as np
import libwave_b200 as W
from libwave_b200 import synth

exe = "tools/native/_build/multimatcher_bench"
for n in (200_000, 1_000_000):
    src, tgt = synth.scan_pair(n)
    src.astype(np.float32).tofile("/tmp/src.f32"); tgt.astype(np.float32).tofile("/tmp/tgt.f32")
    xs, xt = synth.to_xyzw(src), synth.to_xyzw(tgt)
    jobs = 40 if n > 500_000 else 120
    for workers in (1, 2, 4):
        r = subprocess.run([exe, "/tmp/src.f32", "/tmp/tgt.f32", str(workers), str(jobs)], capture_output=True, text=True, timeout=600)
        print("native:", r.stdout.strip() or r.stderr[-300:], flush=True)
        ms_ = [W.ICPMatcher(W.ICPMatcherParams(res=-1)) for _ in range(workers)]
        def work(i, count):
            m = ms_[i]
            for _ in range(count):
                m.setRef(xs); m.setTarget(xt); m.match(); m.estimateInfo()
        def round_of(count):
            th = [threading.Thread(target=work, args=(i, count)) for i in range(workers)]
            [t.start() for t in th]; [t.join() for t in th]
        round_of(2)
        t0 = time.perf_counter(); round_of(jobs // workers); dt = time.perf_counter() - t0
        print(f"python: workers={workers} jobs={jobs // workers * workers} points={n}: {1e3 * dt / (jobs // workers * workers):.3f} ms per job (wall clock), iterations {ms_[0].iterations}", flush=True)
        del ms_
