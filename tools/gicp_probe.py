"""Scratch driver: GICP GPU vs oracle - covariance bit-equality and trajectory agreement per size."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import libwave_b200 as W
from libwave_b200 import synth
from oracle import oracle as O

sizes = [int(a) for a in sys.argv[1:]] or [10_000, 200_000]
for n in sizes:
    src, tgt = synth.scan_pair(n)
    m = W.GICPMatcher(W.GICPMatcherParams(res=-1))
    m.setup(src, tgt)
    cloud, covs = m.covariances(0)
    ref = O.gicp_covariances(src, 10, 1e-3)
    neq = np.any(covs != ref, axis=(1, 2))
    diff = np.abs(covs - ref).max(axis=(1, 2))
    print(f"n={n}: covariances not bit-equal: {int(neq.sum())} of {len(covs)}; max diff {diff.max():.3e}; >1e-9: {int((diff > 1e-9).sum())}")
    if neq.any():
        i = int(np.argmax(diff))
        print("  worst point", i, src[i], "\n  gpu", covs[i].ravel(), "\n  ref", ref[i].ravel())
    t0 = time.time(); ok = m.match(); tg = time.time() - t0
    t0 = time.time(); r = O.gicp_align(src, tgt); to = time.time() - t0
    st = m.stats()
    Tg = m.getResult()
    err = np.abs(Tg[:3, 3] - synth.T_TRUE[:3, 3]).max()
    print(f"  gpu ok={ok} iters={m.iterations} evals={st['evaluations']} ncorr={st['n_corr']} {tg*1e3:.1f} ms | oracle conv={r.converged} iters={r.iterations} "
          f"evals={r.evaluations} ncorr={r.n_corr} {to:.1f} s | max|T_gpu - T_oracle| {np.abs(Tg - r.T).max():.3e} | translation error vs truth {err:.4f} m")
