"""Randomized exactness fuzz of the GPU ICP path against the oracle (tools/fuzz_icp.py with fixed seeds): random
clouds of five shapes (uniform, rings, clusters with duplicates, plane + line, huge extent with a dense core), 1 to
6000 points, NaN / inf injection, max-correspondence distances from 5 cm to 100 m and 1 to 30 iterations, in both
search modes.  Plus the edge the fuzz found: a first iteration with fewer than three pairs (PCL stops with
NO_CORRESPONDENCES, iteration count 0) still reports the pairs it found - pcl::IterativeClosestPoint leaves them in
correspondences_ (reference: the matchers read icp.correspondences_ through the information-matrix estimators,
wave_matching/src/icp.cpp:213)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))


@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz_icp_exact(seed):
    import fuzz_icp
    assert fuzz_icp.run(seed, 60, verbose=True) == 0


def test_fewer_than_three_pairs_reports_them(oracle):
    import libwave_b200 as W
    tgt = np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0], [0, 0, 10], [5, 5, 5]], np.float32)
    src = np.array([[0.1, 0, 0], [50, 50, 50], [-60, 0, 0]], np.float32)
    ref = oracle.icp_align(src, tgt, max_corr=3.0, max_iter=5, sum_mode=oracle.SUM_EXACT)
    assert not ref.converged and ref.iterations == 0 and len(ref.corr_query) == 1
    for mode in (W.SEARCH_TREE, W.SEARCH_TILED):
        m = W.ICPMatcher(W.ICPMatcherParams(res=-1, max_corr=3.0, max_iter=5))
        m.set_search(mode)
        m.setup(src, tgt)
        assert m.match() is False and m.iterations == 0
        q, mm, d2 = m.correspondences()
        assert np.array_equal(q, ref.corr_query) and np.array_equal(mm, ref.corr_match)
        assert np.array_equal(d2, ref.corr_dist)


def test_fuzz_gicp_exact():
    """GICP: both sides add the cost and gradient terms exactly, so flag, outer iterations, evaluation count,
    correspondence count and the final transform agree bit for bit on every random case."""
    import fuzz_gicp_ndt
    assert fuzz_gicp_ndt.run_gicp(21, 50, verbose=True) == 0


def test_fuzz_ndt_tolerance():
    """NDT: flag equal and transform within 1e-4 m / 1e-5 rad with equal iteration counts; the "soft" cases (a step
    or two more or fewer at the 1e-8 stop threshold, or an oracle path that ran into the iteration cap) are
    counted, not failed (tools/fuzz_gicp_ndt.py run_ndt)."""
    import fuzz_gicp_ndt
    bad, soft = fuzz_gicp_ndt.run_ndt(22, 60, verbose=True)
    assert bad == 0
    assert soft <= 6


def test_fuzz_matcher_branches_exact():
    """ICPMatcher::match() in its three branches (full resolution, voxel grid, multiscale; src/icp.cpp:75-133) plus
    estimateLUM / estimateLUMold on random scenes: flag, iterations, transform, correspondences, aligned cloud and
    both information matrices bit-equal to the oracle's."""
    import fuzz_match
    assert fuzz_match.run(31, 50, verbose=True) == 0
