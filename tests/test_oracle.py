"""CPU tests (-m "not gpu"): the oracle against brute force, against numpy/scipy, and against what
the reference's own tests and the survey pin for this path."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from conftest import pcl_transform, rot_angle


def test_kdtree_matches_bruteforce_random(oracle):
    rng = np.random.default_rng(0)
    for n_t, n_q in ((1, 50), (15, 100), (16, 100), (3000, 2000)):
        tgt = rng.normal(size=(n_t, 3)).astype(np.float32)
        qry = rng.normal(size=(n_q, 3)).astype(np.float32)
        i1, d1 = oracle.KdTree(tgt).nn1(qry)
        i2, d2 = oracle.brute_nn1(tgt, qry)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2)


def test_kdtree_ties_pick_lowest_index(oracle):
    tgt = np.array([[1, 0, 0], [0, 1, 0], [1, 0, 0], [-1, 0, 0]], np.float32)
    idx, d2 = oracle.KdTree(tgt).nn1(np.zeros((1, 3), np.float32))
    assert idx[0] == 0 and d2[0] == 1.0
    many = np.tile(np.array([[0.5, 0.25, -1.0]], np.float32), (100, 1))  # 100 duplicates
    idx, _ = oracle.KdTree(many).nn1(np.array([[0.4, 0.2, -1.0]], np.float32))
    assert idx[0] == 0


def test_kdtree_skips_non_finite_targets(oracle):
    tgt = np.array([[0, 0, 0], [np.nan, 0, 0], [5, 5, 5], [np.inf, 1, 1]], np.float32)
    idx, _ = oracle.KdTree(tgt).nn1(np.array([[4, 4, 4], [0.1, 0, 0]], np.float32))
    assert idx.tolist() == [2, 0]
    empty_idx, _ = oracle.KdTree(np.zeros((0, 3), np.float32)).nn1(np.zeros((2, 3), np.float32))
    assert (empty_idx == -1).all()


def test_kdtree_testscan_vs_scipy_and_tie_count(oracle, testscan):
    """one duplicated point -> exactly 2 queries with an exact distance tie;
    the fp32 argmin equals the fp64 argmin everywhere else."""
    T = np.eye(4)
    T[0, 3] = 0.2
    tgt = pcl_transform(testscan, T)
    idx, d2 = oracle.KdTree(tgt).nn1(testscan, nthreads=4)
    dd, ii = cKDTree(tgt.astype(np.float64)).query(testscan.astype(np.float64), k=2)
    ties = np.sum(dd[:, 0] == dd[:, 1])
    assert ties == 2
    agree = (ii[:, 0] == idx) | (dd[:, 0] == dd[:, 1])
    assert agree.all()
    assert np.allclose(np.sqrt(d2.astype(np.float64)), dd[:, 0], rtol=1e-5, atol=1e-7)


def test_knn_sorted_and_consistent(oracle):
    rng = np.random.default_rng(2)
    tgt = rng.uniform(-1, 1, size=(500, 3)).astype(np.float32)
    qry = rng.uniform(-1, 1, size=(40, 3)).astype(np.float32)
    idx, d2 = oracle.KdTree(tgt).knn(qry, 10)
    assert (np.diff(d2, axis=1) >= 0).all()
    full = ((qry[:, None, :].astype(np.float64) - tgt[None].astype(np.float64)) ** 2).sum(-1)
    assert np.array_equal(np.sort(np.argsort(full, axis=1)[:, :10], axis=1), np.sort(idx, axis=1))


def test_rotation_from_sigma_matches_numpy_svd(oracle):
    rng = np.random.default_rng(5)
    for _ in range(50):
        S = rng.normal(size=(3, 3))
        U, _, Vt = np.linalg.svd(S)
        D = np.diag([1, 1, np.sign(np.linalg.det(U) * np.linalg.det(Vt))])
        assert np.allclose(oracle.rotation_from_sigma(S), U @ D @ Vt, atol=1e-10)
    flat = np.outer([1, 2, 3], [0.5, -1, 2]) + np.outer([0, 1, -1], [1, 1, 0])  # rank 2
    R = oracle.rotation_from_sigma(flat)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-10) and np.isclose(np.linalg.det(R), 1.0)


def test_solve6_matches_numpy(oracle):
    rng = np.random.default_rng(6)
    A = rng.normal(size=(6, 6))
    A = A @ A.T + np.eye(6)
    b = rng.normal(size=6)
    assert np.allclose(oracle.solve6(A, b), np.linalg.solve(A, b), rtol=1e-10)
    assert oracle.solve6(np.zeros((6, 6)), b) is None


def test_voxel_grid_counts_match_survey(oracle, testscan):
    """Occupied-voxel counts measured independently in SURVEY.md Appendix B."""
    expect = {0.05: 30024, 0.1: 17598, 0.2: 8642, 0.4: 3569, 0.8: 1322}
    for leaf, n in expect.items():
        out, ok = oracle.voxel_grid(testscan, leaf)
        assert ok and out.shape[0] == n
    out, ok = oracle.voxel_grid(testscan, 0.4)
    # centroids lie inside the bounding box and voxels are emitted in ascending index order
    assert (out[:, :3].min(0) >= testscan.min(0) - 1e-6).all() and (out[:, :3].max(0) <= testscan.max(0) + 1e-6).all()
    out, ok = oracle.voxel_grid(testscan[:2000], 1e-4)
    assert not ok and np.array_equal(out[:, :3], testscan[:2000])  # overflow rule: input passed through
    out, ok = oracle.voxel_grid(np.zeros((0, 3), np.float32), 0.1)
    assert out.shape[0] == 0


@pytest.mark.parametrize("sum_mode", [0, 1])
def test_reference_icp_tests_restated(oracle, testscan, sum_mode):
    """tests/icp_tests.cpp:45-148 against the oracle: match()==true and ||result - perturb||_F < 0.1,
    plus the iteration counts / stop rules the survey's scratch run reports (Appendix D)."""
    cases = [  # res, multiscale_steps, tx, expected total iterations
        (-1.0, 0, 0.0, 1),      # fullResNullMatch
        (0.05, 0, 0.0, 1),      # nullDisplacement
        (0.05, 0, 0.2, 8),      # smallDisplacement: TRANSFORM rule after 8 iterations
        (0.1, 3, 0.2, None),    # multiscale
    ]
    for res, steps, tx, iters in cases:
        T = np.eye(4)
        T[0, 3] = tx
        tgt = pcl_transform(testscan, T)
        r = oracle.icp_match(testscan, tgt, res=res, multiscale_steps=steps, sum_mode=sum_mode, nn_threads=4)
        assert r.success
        assert np.linalg.norm(r.T - T) < 0.1
        assert np.linalg.norm(r.T - T) < 1e-5  # far inside the reference's bound on clean data
        if iters is not None:
            assert r.total_iterations == iters
    full = oracle.icp_align(testscan, pcl_transform(testscan, T), sum_mode=sum_mode, nn_threads=4)
    assert full.state == "TRANSFORM" and full.iterations == 11  # Appendix D: TRANSFORM, 11
    assert np.isclose(full.mse[0], 1.18e-2, rtol=2e-2)           # Appendix D: first MSE 1.18e-2


def test_exact_and_pcl_arithmetic_agree_within_tolerance(oracle):
    """The estimator spec used by the GPU (exact sums, fp64 Umeyama) against the PCL-faithful fp32
    arithmetic: the gap must stay inside the north-star tolerance (1e-4 m, 1e-5 rad)."""
    from libwave_b200 import synth
    src, tgt, nrm = synth.scan_pair(10_000, return_normals=True)
    for est in (oracle.EST_SVD, oracle.EST_POINT_TO_PLANE):
        a = oracle.icp_align(src, tgt, estimator=est, sum_mode=oracle.SUM_EXACT, target_normals=nrm)
        b = oracle.icp_align(src, tgt, estimator=est, sum_mode=oracle.SUM_PCL, target_normals=nrm)
        assert a.iterations == b.iterations and a.state == b.state
        assert np.abs(a.T[:3, 3] - b.T[:3, 3]).max() < 1e-4
        Ra, Rb = a.T[:3, :3].astype(np.float64), b.T[:3, :3].astype(np.float64)
        assert rot_angle(Ra, Rb) < 1e-5


def test_point_to_plane_converges_closer_than_svd(oracle):
    from libwave_b200 import synth
    src, tgt, nrm = synth.scan_pair(10_000, return_normals=True)
    tgt2 = pcl_transform(src, synth.T_TRUE)  # a rigidly moved copy: the premise of the reference tests
    r = oracle.icp_align(src, tgt2, estimator=oracle.EST_POINT_TO_PLANE, target_normals=nrm, max_iter=5)
    assert r.iterations <= 5 and r.converged  # hits the iteration cap -> PCL still reports converged


def test_degenerate_inputs(oracle):
    empty = np.zeros((0, 3), np.float32)
    pts = np.random.default_rng(1).normal(size=(100, 3)).astype(np.float32)
    assert not oracle.icp_align(empty, pts).converged
    assert not oracle.icp_align(pts, empty).converged
    far = pts + np.float32(100.0)
    r = oracle.icp_align(pts, far, max_corr=1.0)  # no correspondence within max_corr
    assert not r.converged and r.state == "NO_CORRESPONDENCES"
    r = oracle.icp_align(pts, pts, max_iter=1)
    assert r.converged and r.iterations == 1


def test_information_matrix_properties(oracle, testscan):
    rng = np.random.default_rng(0)
    T = np.eye(4)
    T[0, 3] = 0.2
    tgt = (pcl_transform(testscan, T).astype(np.float64) + rng.uniform(-0.3, 0.3, testscan.shape)).astype(np.float32)
    r = oracle.icp_match(testscan, tgt, res=0.05, multiscale_steps=0, nn_threads=4)
    k_quad = oracle.fix_scales(r.ds_ref, r.ds_tgt, 3.0)[1]
    for mode in (oracle.SUM_EXACT, oracle.SUM_PCL):
        lum, ok = oracle.estimate_lum(r.last.aligned, r.ds_tgt, r.last.corr_query, r.last.corr_match, mode, k_quad)
        old, ok2 = oracle.estimate_lum_old(r.last.aligned, r.ds_tgt, 3.0, mode, k_quad, nn_threads=4)
        assert ok and ok2
        for info in (lum, old):
            assert info[0, 0] > 0                      # tests/icp_tests.cpp:123,193
            assert np.allclose(info, info.T)
            assert np.all(np.linalg.eigvalsh(info) > 0)
    a, _ = oracle.estimate_lum_old(r.last.aligned, r.ds_tgt, 3.0, oracle.SUM_EXACT, k_quad, nn_threads=4)
    b, _ = oracle.estimate_lum_old(r.last.aligned, r.ds_tgt, 3.0, oracle.SUM_PCL, k_quad, nn_threads=4)
    assert np.allclose(a, b, rtol=1e-4)


def test_fix128_small_negative_sums_convert_exactly(oracle):
    """The 128-bit totals are converted sign-magnitude: a small negative total keeps its low bits
    (two's-complement halves would round 2^64 - |v| to 53 bits first and return 0 for -5 units)."""
    assert oracle.fix128_sum([-5.0], 0) == -5.0
    assert oracle.fix128_sum([3.0, -8.0], 0) == -5.0
    assert oracle.fix128_sum([-5.0 * 2.0 ** -47], 47) == -5.0 * 2.0 ** -47
    big = 2.0 ** 40
    assert oracle.fix128_sum([big, -big, -1.25], 20) == -1.25
    rng = np.random.default_rng(5)
    t = rng.normal(0, 50.0, 100_000)
    k = 30
    exact = sum(int(np.rint(np.ldexp(v, k))) for v in t)      # Python integers: exact
    assert oracle.fix128_sum(t, k) == float(exact) / 2.0 ** k
    assert oracle.fix128_sum(-np.abs(t), k) == -oracle.fix128_sum(np.abs(t), k)
