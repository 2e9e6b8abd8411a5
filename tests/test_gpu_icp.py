"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C ABI against the
CPU oracle on the same inputs.  Bars: correspondence indices and fp32 squared distances bit-exact;
per-iteration transforms, iteration count, stop rule and final transform bit-exact against the
oracle's SUM_EXACT arithmetic (the repo's estimator spec); within 1e-4 m / 1e-5 rad of
the oracle's PCL-faithful fp32 arithmetic (BASELINE.json north_star tolerance)."""
import numpy as np
import pytest

from conftest import pcl_transform

pytestmark = pytest.mark.gpu


def rot_angle(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1.0) / 2.0
    return float(np.arccos(np.clip(c, -1.0, 1.0)))


@pytest.fixture(scope="module")
def W():
    import libwave_b200 as W
    return W


def test_nn_matches_bruteforce_small(W, oracle):
    rng = np.random.default_rng(7)
    tgt = rng.normal(size=(5000, 3)).astype(np.float32)
    qry = rng.normal(size=(3000, 3)).astype(np.float32)
    nn = W.NearestNeighbour(tgt)
    idx, d2 = nn.search(qry)
    ridx, rd2 = oracle.brute_nn1(tgt, qry)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(d2, rd2)


def test_nn_matches_oracle_testscan(W, oracle, testscan):
    T = np.eye(4)
    T[0, 3] = 0.2
    tgt = pcl_transform(testscan, T)
    nn = W.NearestNeighbour(tgt)
    idx, d2 = nn.search(testscan)
    ridx, rd2 = oracle.KdTree(tgt).nn1(testscan, nthreads=8)
    assert np.array_equal(d2, rd2)
    assert np.array_equal(idx, ridx)  # includes the two exact-tie queries (lowest index rule)


def test_nn_max_dist_and_edge_cases(W, oracle):
    rng = np.random.default_rng(3)
    tgt = rng.uniform(-5, 5, size=(2000, 3)).astype(np.float32)
    qry = rng.uniform(-8, 8, size=(1500, 3)).astype(np.float32)
    qry[5] = np.nan
    tgt[7] = np.inf
    nn = W.NearestNeighbour(tgt)
    idx, d2 = nn.search(qry, max_dist=1.0)
    ridx, rd2 = oracle.brute_nn1(tgt, qry[np.isfinite(qry).all(1)])
    keep = np.isfinite(qry).all(1)
    ri = np.full(len(qry), -1, np.int32)
    rd = np.full(len(qry), np.inf, np.float32)
    ri[keep], rd[keep] = ridx, rd2
    far = rd.astype(np.float64) > 1.0
    ri[far], rd[far] = -1, np.inf
    assert np.array_equal(idx, ri)
    assert np.array_equal(d2, rd)
    # single-point and empty targets
    one = W.NearestNeighbour(tgt[:1])
    i1, _ = one.search(qry[:10])
    assert (i1[np.isfinite(qry[:10]).all(1)] == 0).all()
    empty = W.NearestNeighbour(np.zeros((0, 3), np.float32))
    i0, _ = empty.search(qry[:10])
    assert (i0 == -1).all()


@pytest.mark.parametrize("tx", [0.0, 0.2])
def test_icp_fullres_bit_exact_vs_oracle(W, oracle, testscan, tx):
    """fullResNullMatch (tests/icp_tests.cpp:45-62) and the 0.2 m displacement at full resolution."""
    T = np.eye(4)
    T[0, 3] = tx
    tgt = pcl_transform(testscan, T)
    prm = W.ICPMatcherParams(res=-1)
    m = W.ICPMatcher(prm)
    m.setup(testscan, tgt)
    assert m.match() is True
    assert np.linalg.norm(m.getResult() - T) < 0.1  # the reference test's own bound
    ref = oracle.icp_align(testscan, tgt, sum_mode=oracle.SUM_EXACT, nn_threads=8)
    mse, ncorr, Ttr = m.trace()
    assert m.iterations == ref.iterations
    assert np.array_equal(ncorr, ref.n_corr)
    assert np.array_equal(Ttr, ref.T_trace)
    assert np.array_equal(mse, ref.mse)
    assert np.array_equal(m.getResult().astype(np.float32), ref.T)
    q, mm, d2 = m.correspondences()
    assert np.array_equal(q, ref.corr_query)
    assert np.array_equal(mm, ref.corr_match)
    assert np.array_equal(d2, ref.corr_dist)
    assert np.array_equal(m.aligned()[:, :3], ref.aligned[:, :3])
    # noise floor of the reference's own fp32 arithmetic
    pcl = oracle.icp_align(testscan, tgt, sum_mode=oracle.SUM_PCL, nn_threads=8)
    assert np.abs(m.getResult()[:3, 3] - pcl.T[:3, 3].astype(np.float64)).max() < 1e-4
    assert rot_angle(m.getResult()[:3, :3], pcl.T[:3, :3].astype(np.float64)) < 1e-5
