"""Randomized exactness fuzz of ICPMatcher::match() (src/icp.cpp:75-133: full resolution / voxel grid / multiscale
branches) and the LUM / LUMold information estimators against the oracle: success flag, total iteration count,
final transform, correspondences, aligned cloud and the two information matrices must agree bit for bit.
`python tools/fuzz_match.py SEED N_CASES`; run with fixed seeds by tests/test_gpu_fuzz.py."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from fuzz_gicp_ndt import scene, rot


def run(seed, n_cases, verbose=True):
    import libwave_b200 as W
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    bad = 0
    for case in range(n_cases):
        n = int(rng.choice([40, 300, 2000, 8000]))
        tgt = scene(rng, int(rng.integers(0, 5)), n)
        t = rng.normal(0, rng.choice([0.0, 0.05, 0.3]), 3); a = rng.normal(0, rng.choice([0.0, 0.01, 0.05]), 3)
        src = ((tgt.astype(np.float64) - t) @ rot(*a) + rng.normal(0, 0.005, tgt.shape)).astype(np.float32)
        if rng.random() < 0.3:
            src = src[rng.permutation(n)[: max(4, int(n * rng.uniform(0.4, 1.0)))]]
        if rng.random() < 0.15:
            src[rng.integers(0, len(src), max(1, len(src) // 20))] = np.nan
        kw = dict(res=float(rng.choice([-1.0, 0.05, 0.2, 1.0])), multiscale_steps=int(rng.choice([0, 1, 3])),
                  max_corr=float(rng.choice([0.5, 3.0])), max_iter=int(rng.choice([2, 10, 100])))
        ref = O.icp_match(src, tgt, nn_threads=4, **kw)
        m = W.ICPMatcher(W.ICPMatcherParams(**kw))
        m.setup(src, tgt)
        ok = m.match()
        what = []
        if ok != ref.success: what.append("flag")
        if m.iterations != ref.total_iterations: what.append(f"iterations {m.iterations} vs {ref.total_iterations}")
        if not np.array_equal(m.getResult(), ref.T, equal_nan=True): what.append("T")
        q, mm, d2 = m.correspondences()
        if not (np.array_equal(q, ref.last.corr_query) and np.array_equal(mm, ref.last.corr_match)
                and np.array_equal(d2, ref.last.corr_dist)): what.append(f"correspondences {len(q)} vs {len(ref.last.corr_query)}")
        if not np.array_equal(m.aligned()[:, :3], ref.last.aligned[:, :3], equal_nan=True): what.append("aligned")
        if not what and len(q) >= 6:
            k_quad = O.fix_scales(ref.ds_ref, ref.ds_tgt, kw["max_corr"])[1]
            r_lum, ok_lum = O.estimate_lum(ref.last.aligned, ref.ds_tgt, ref.last.corr_query, ref.last.corr_match, O.SUM_EXACT, k_quad)
            r_old, ok_old = O.estimate_lum_old(ref.last.aligned, ref.ds_tgt, kw["max_corr"], O.SUM_EXACT, k_quad, nn_threads=4)
            if ok_lum and not np.array_equal(m.info(W.INFO_LUM), r_lum, equal_nan=True): what.append("LUM")
            if ok_old and not np.array_equal(m.info(W.INFO_LUMOLD), r_old, equal_nan=True): what.append("LUMold")
        if what:
            bad += 1
            if verbose:
                print(f"MATCH MISMATCH case {case}: n_s={len(src)} n_t={n} {kw} levels={ref.levels}: {', '.join(what)}")
            if bad > 5:
                break
    return bad


if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    b = run(seed, n)
    print(f"fuzz match: {n} cases, {b} mismatches")
    sys.exit(1 if b else 0)
