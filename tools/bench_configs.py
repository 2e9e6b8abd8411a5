#!/usr/bin/env python
"""Timings of the BASELINE.json configurations that are not the headline bench line (scratch tool;
its output is kept under profiles/):

  C3  GICPMatcher, per-point covariances, 500k-point scan pair          (configs[2])
  C4  NDTMatcher, 0.5 m voxels, 1M-point scan vs 5M-point map           (configs[3])

For each: wall time of match() from host clouds (uploads and result read-back inside), the number
of full passes over the source the optimiser asked for (one "pair" = one source point in one
cost / derivative evaluation), and the CPU oracle on a bounded sample of the same
generator (smaller clouds, same parameters), single-threaded as PCL's optimisers are.

    python tools/bench_configs.py [--cpu]
"""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np

import libwave_b200 as W
from libwave_b200 import synth


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), out


def gicp(cpu: bool):
    n = 500_000
    src, tgt = (synth.to_xyzw(a) for a in synth.scan_pair(n))   # xyzw up front: time the library, not numpy
    m = W.GICPMatcher(W.GICPMatcherParams(res=-1))

    def run():
        m.setup(src, tgt)
        return m.match()
    dt, ok = timed(run)
    st = m.stats()
    line = {"config": "GICPMatcher with per-point covariance, 500k-point scan pair, 1xB200", "converged": bool(ok),
            "outer_iterations": m.iterations, "cost_gradient_evaluations": st["evaluations"],
            "correspondences": st["n_corr"], "kernel_launches": st["kernel_launches"], "match_ms": 1e3 * dt,
            "point_pairs_per_s": st["evaluations"] * n / dt,
            "translation_error_m": float(np.abs(m.getResult()[:3, 3] - synth.T_TRUE[:3, 3]).max())}
    if cpu:
        from oracle import oracle as O
        s2, t2 = synth.scan_pair(10_000)
        t0 = time.perf_counter()
        r = O.gicp_align(s2, t2)
        dc = time.perf_counter() - t0
        line["cpu_oracle"] = {"sample": f"{len(s2)}-point pair, same generator and parameters, 1 thread",
                              "seconds": dc, "evaluations": int(r.evaluations),
                              "point_pairs_per_s": r.evaluations * len(s2) / dc}
    return line


def ndt(cpu: bool):
    rings, az = synth.SIZES[1_000_000]
    scan = synth.velodyne_scan(rings, az, None, synth.SOURCE_SEED, n_points=1_000_000)
    big = synth.map_cloud(5, 1_000_000)
    scan, big = synth.to_xyzw(scan), synth.to_xyzw(big)
    m = W.NDTMatcher(W.NDTMatcherParams(res=0.5))

    def run():
        m.setup(scan, big)
        return m.match()
    dt, ok = timed(run, reps=2)
    st = m.stats()
    line = {"config": "NDTMatcher 0.5 m voxel grid, 1M-point scan vs 5M-point map, 1xB200", "converged": bool(ok),
            "iterations": m.iterations, "derivative_passes": st["derivative_passes"], "cells": st["n_cells"],
            "map_points": int(len(big)), "kernel_launches": st["kernel_launches"], "match_ms": 1e3 * dt,
            "point_pairs_per_s": st["derivative_passes"] * len(scan) / dt}
    if cpu:
        from oracle import oracle as O
        s2 = scan[:: 20]
        b2 = big[:: 20]
        t0 = time.perf_counter()
        r = O.ndt_align(s2, b2, res=0.5, max_iter=5)
        dc = time.perf_counter() - t0
        line["cpu_oracle"] = {"sample": f"every 20th point ({len(s2)} vs {len(b2)}), max_iter 5, 1 thread",
                              "seconds": dc, "iterations": int(r.iterations)}
    return line


if __name__ == "__main__":
    cpu = "--cpu" in sys.argv
    for fn in (gicp, ndt):
        print(json.dumps(fn(cpu)), flush=True)
