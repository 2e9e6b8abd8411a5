"""GPU parity tests of the tiled correspondence kernel (csrc/tile_nn.cuh, WAVECU_SEARCH_TILED): it
must reproduce the tree walk - and therefore the oracle - bit for bit: iteration count, MSE trace,
per-iteration transforms, correspondence indices and fp32 distances.  Covered: the reference fixture,
tiny / duplicated targets, degenerate geometry (tiles that cannot be staged and fall back to the
walk entirely), non-finite points, max-distance rejections, the multiscale branch, and the 200 k /
1 M BASELINE sizes."""
import numpy as np
import pytest

from conftest import pcl_transform

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def W():
    import libwave_b200 as W
    return W


def run(W, src, tgt, mode, nrm=None, **kw):
    prm = W.ICPMatcherParams(res=kw.pop("res", -1), **kw)
    m = W.ICPMatcher(prm)
    m.set_search(mode)
    m.setup(src, tgt)
    if nrm is not None:
        m.setTargetNormals(nrm)
    ok = m.match()
    return m, ok


def assert_same(W, src, tgt, nrm=None, oracle=None, **kw):
    a, oka = run(W, src, tgt, W.SEARCH_TREE, nrm, **kw)
    b, okb = run(W, src, tgt, W.SEARCH_TILED, nrm, **kw)
    assert oka == okb and a.iterations == b.iterations
    ta, tb = a.trace(), b.trace()
    for u, v in zip(ta, tb):
        assert np.array_equal(u, v)
    assert np.array_equal(a.getResult(), b.getResult())
    qa, ma, da = a.correspondences()
    qb, mb, db = b.correspondences()
    assert np.array_equal(qa, qb) and np.array_equal(ma, mb) and np.array_equal(da, db)
    st = b.stats()
    assert 0 <= st["fallback_queries"] <= st["pairs"]
    return a, b


def test_tiled_reference_fixture_equals_tree_and_oracle(W, oracle, testscan):
    T = np.eye(4)
    T[0, 3] = 0.2
    tgt = pcl_transform(testscan, T)
    a, b = assert_same(W, testscan, tgt)
    ref = oracle.icp_align(testscan, tgt, sum_mode=oracle.SUM_EXACT, nn_threads=8)
    q, mm, d2 = b.correspondences()
    assert b.iterations == ref.iterations
    assert np.array_equal(mm, ref.corr_match) and np.array_equal(d2, ref.corr_dist)
    assert np.array_equal(b.getResult().astype(np.float32), ref.T)
    assert np.array_equal(b.aligned()[:, :3], ref.aligned[:, :3])


@pytest.mark.parametrize("res,steps", [(0.05, 0), (0.1, 3)])
def test_tiled_voxel_and_multiscale_branches(W, testscan, res, steps):
    T = np.eye(4)
    T[0, 3] = 0.2
    assert_same(W, testscan, pcl_transform(testscan, T), res=res, multiscale_steps=steps)


@pytest.mark.parametrize("n_tgt", [1, 3, 9, 33, 64, 65, 513])
def test_tiled_tiny_targets_duplicates_and_ties(W, n_tgt):
    rng = np.random.default_rng(n_tgt)
    base = rng.uniform(-2, 2, size=(n_tgt, 3)).astype(np.float32)
    tgt = np.concatenate([base, base[::-1]])             # every point twice: exact ties everywhere
    qry = rng.uniform(-3, 3, size=(700, 3)).astype(np.float32)
    qry[:n_tgt] = base[:700]
    if 2 * n_tgt >= 3:
        assert_same(W, qry, tgt, max_corr=100.0, max_iter=3)


@pytest.mark.parametrize("shape", ["slab", "line", "clusters", "far_queries", "nonfinite", "tight_max_corr"])
def test_tiled_adversarial_geometry(W, shape):
    rng = np.random.default_rng({"slab": 1, "line": 2, "clusters": 3, "far_queries": 4, "nonfinite": 5,
                                 "tight_max_corr": 6}[shape])
    n = 6000
    if shape == "slab":
        tgt = rng.uniform([-100, -75, -0.01], [100, 75, 0.01], size=(n, 3))
    elif shape == "line":
        tgt = np.zeros((n, 3))
        tgt[:, 0] = rng.uniform(-50, 50, n)
    elif shape == "clusters":
        centers = rng.uniform(-60, 60, size=(12, 3))
        tgt = centers[rng.integers(0, 12, n)] + rng.normal(0, 0.02, (n, 3))
        tgt[n // 2:] = tgt[: n - n // 2]
    else:
        tgt = rng.uniform(-5, 5, size=(n, 3))
    tgt = tgt.astype(np.float32)
    src = (tgt[rng.permutation(n)] + rng.normal(0, 0.3, (n, 3))).astype(np.float32)
    kw = dict(max_corr=5.0, max_iter=5)
    if shape == "far_queries":
        src[::7] += np.float32(40.0)
        kw["max_corr"] = 100.0
    if shape == "nonfinite":
        src[::13] = np.nan
        tgt[::17] = np.inf
        tgt[5] = np.nan
    if shape == "tight_max_corr":
        kw["max_corr"] = 0.25                    # most pairs rejected by the distance test
    src[::11] += rng.uniform(-2.5, 2.5, (len(src[::11]), 3)).astype(np.float32)
    assert_same(W, src, tgt, **kw)


def test_tiled_point_to_plane_200k(W, oracle):
    from libwave_b200 import synth
    src, tgt, nrm = synth.scan_pair(200_000, return_normals=True)
    a, b = assert_same(W, src, tgt, nrm, estimator=W.EST_POINT_TO_PLANE)
    ref = oracle.icp_align(src, tgt, estimator=oracle.EST_POINT_TO_PLANE, sum_mode=oracle.SUM_EXACT,
                           target_normals=nrm, nn_threads=8)
    q, mm, d2 = b.correspondences()
    assert b.iterations == ref.iterations
    assert np.array_equal(mm, ref.corr_match) and np.array_equal(d2, ref.corr_dist)


def test_tiled_1m_svd(W):
    from libwave_b200 import synth
    src, tgt = synth.scan_pair(1_000_000)
    a, b = assert_same(W, src, tgt)
    assert b.stats()["fallback_queries"] < 0.5 * b.stats()["pairs"]
