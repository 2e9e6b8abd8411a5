"""Randomized exactness fuzz - GPU ICP (tree and tiled search; WAVECU_FUSED=0/1 picks unfused / fused iterations)
against the oracle on random clouds of assorted shapes, sizes and parameters: converged flag, iteration count, the
correspondence arrays (bit for bit) and the transform must agree.  `python tools/fuzz_icp.py SEED N_CASES`; also run
with a fixed seed by tests/test_gpu_fuzz.py."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np


def cloud(rng, kind, n):
    if kind == 0:
        p = rng.uniform(-20, 20, (n, 3))
    elif kind == 1:   # rings: dense along, sparse across
        a = rng.uniform(0, 2 * np.pi, n); r = rng.choice(np.linspace(3, 40, 12), n)
        p = np.stack([r * np.cos(a), r * np.sin(a), rng.normal(0, 0.01, n) - 1.7], 1)
    elif kind == 2:   # clusters with duplicates
        c = rng.uniform(-30, 30, (8, 3)); p = c[rng.integers(0, 8, n)] + rng.normal(0, 0.05, (n, 3))
        p[n // 2:] = p[: n - n // 2]
    elif kind == 3:   # plane + line
        p = np.zeros((n, 3)); p[:, 0] = rng.uniform(-50, 50, n); p[: n // 2, 1] = rng.uniform(-50, 50, n // 2)
    else:             # huge extent with a dense core
        p = np.concatenate([rng.normal(0, 0.3, (n // 2, 3)), rng.uniform(-500, 500, (n - n // 2, 3))])
    return p.astype(np.float32)


def run(seed, n_cases, verbose=True):
    import libwave_b200 as W
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    bad = 0
    for case in range(n_cases):
        n_t = int(rng.choice([1, 2, 5, 17, 63, 64, 65, 300, 2000, 6000])); n_s = int(rng.choice([1, 3, 33, 257, 1000, 5000]))
        tgt = cloud(rng, int(rng.integers(0, 5)), n_t)
        src = (tgt[rng.integers(0, n_t, n_s)] + rng.normal(0, rng.choice([0.01, 0.3, 3.0]), (n_s, 3))).astype(np.float32)
        if rng.random() < 0.2:
            src[rng.integers(0, n_s, max(1, n_s // 10))] = np.nan
        if rng.random() < 0.2:
            tgt[rng.integers(0, n_t, max(1, n_t // 10))] = np.inf
        kw = dict(max_corr=float(rng.choice([0.05, 0.5, 3.0, 100.0])), max_iter=int(rng.choice([1, 2, 5, 30])))
        ref = O.icp_align(src, tgt, sum_mode=O.SUM_EXACT, **kw)
        for mode in (W.SEARCH_TREE, W.SEARCH_TILED):
            m = W.ICPMatcher(W.ICPMatcherParams(res=-1, **kw))
            m.set_search(mode)
            m.setup(src, tgt)
            ok = m.match()
            q, mm, d2 = m.correspondences()
            same = (ok == ref.converged and m.iterations == ref.iterations and np.array_equal(q, ref.corr_query)
                    and np.array_equal(mm, ref.corr_match) and np.array_equal(d2, ref.corr_dist)
                    and (not ref.converged or np.array_equal(m.getResult().astype(np.float32), ref.T)))
            if not same:
                bad += 1
                if verbose: print(f"MISMATCH case {case} mode {mode}: n_s={n_s} n_t={n_t} {kw} gpu(ok={ok}, it={m.iterations}, nc={len(q)}) "
                      f"oracle(ok={ref.converged}, it={ref.iterations}, nc={len(ref.corr_query)}, state={ref.state})")
                if bad > 5:
                    return bad
    return bad


if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    bad = run(seed, n)
    print(f"fuzz: {n} cases x 2 search modes, {bad} mismatches (WAVECU_FUSED={os.environ.get('WAVECU_FUSED', '1')})")
    sys.exit(1 if bad else 0)
