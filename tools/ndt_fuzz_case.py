"""Scratch: replay one case of tools/fuzz_gicp_ndt.py run_ndt (seed, case index) and print both step traces."""
import os, sys
os.environ["WAVECU_NDT_TRACE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import fuzz_gicp_ndt as F
import libwave_b200 as W
from oracle import oracle as O
seed, want = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(seed)
for case in range(want + 1):
    src, tgt = F.make_pair(rng)
    kw = dict(step_size=int(rng.choice([1, 3])), max_iter=int(rng.choice([1, 5, 35])), res=float(rng.choice([0.5, 1.0, 3.0])),
              line_search=int(rng.integers(0, 2)))
print(kw, len(src), len(tgt))
ref = O.ndt_align(src, tgt, **kw)
print("oracle steps", [float(f"{v:.6g}") for v in ref.steps], ref.iterations, ref.score)
m = W.NDTMatcher(W.NDTMatcherParams(**kw))
m.setup(src, tgt)
print(m.match(), m.iterations)
print("dT", np.abs(m.getResult() - ref.T).max())
print(m.getResult()); print(ref.T)
print("nonfinite src", (~np.isfinite(src).all(1)).sum(), "tgt", (~np.isfinite(tgt).all(1)).sum())
voxel, count, cen, mean, icov = m.grid()
rv, rc, rcen, rmean, ricov = O.ndt_grid(tgt, kw["res"])
print("grid", len(voxel), len(rv), np.array_equal(voxel, rv), np.array_equal(count, rc), np.array_equal(cen, rcen))
if len(voxel) == len(rv):
    print(" mean", np.abs(mean - rmean).max(), "icov rel", (np.abs(icov - ricov).max(axis=(1, 2)) / np.abs(ricov).max(axis=(1, 2))).max())
p0 = [0, 0, 0, 0, 0, 0]
s, g, H = m.derivatives(p0, np.eye(4, dtype=np.float32))
rs, rg, rH = O.ndt_derivatives(src, tgt, kw["res"], p0, np.eye(4, dtype=np.float32))
print("score", s, rs, "g", np.abs(g - rg).max() / np.abs(rg).max(), "H", np.abs(H - rH).max() / np.abs(rH).max())
only_gpu = np.setdiff1d(voxel, rv); only_ref = np.setdiff1d(rv, voxel)
print("only gpu", only_gpu, "only oracle", only_ref)
for v in only_gpu:
    i = int(np.where(voxel == v)[0][0])
    print(" gpu cell", v, "count", count[i], "mean", mean[i], "icov", icov[i].ravel())
    # points of that voxel
    lo = np.floor(tgt.min(0) / kw["res"]); ijk = np.floor(tgt / kw["res"]) - lo
    print(" z spread of tgt", tgt[:, 2].min(), tgt[:, 2].max())
