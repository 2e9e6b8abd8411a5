"""GPU tests at BASELINE.json's full sizes, through size-independent properties (the oracle would
need minutes per case at these sizes):
  * a rigidly moved copy of a cloud (the premise of every reference test) must be recovered, and at
    convergence every source point must correspond to its own copy - an index-exact property;
  * nearest neighbours of a cloud in itself are the identity permutation with zero distance;
  * results do not depend on the order of the input points (the estimator sums are exact);
  * the NDT score at the true pose beats the score at identity on the 1M-vs-5M configuration."""
import numpy as np
import pytest

from conftest import pcl_transform, rot_angle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def W():
    import libwave_b200 as W
    return W


@pytest.fixture(scope="module")
def synth():
    from libwave_b200 import synth
    return synth


@pytest.fixture(scope="module")
def scan_1m(synth):
    rings, az = synth.SIZES[1_000_000]
    return synth.velodyne_scan(rings, az, None, synth.SOURCE_SEED, n_points=1_000_000)


def test_nn_1m_self_is_identity(W, scan_1m):
    nn = W.NearestNeighbour(scan_1m)
    idx, d2 = nn.search(scan_1m)
    assert np.array_equal(idx, np.arange(len(scan_1m), dtype=np.int32))   # the generator de-duplicates
    assert not d2.any()


def test_icp_1m_recovers_rigid_copy_index_exact(W, synth, scan_1m):
    """BASELINE config 2 size, reference-test premise: target = T_true * source (same point order)."""
    tgt = pcl_transform(scan_1m, synth.T_TRUE)
    m = W.ICPMatcher(W.ICPMatcherParams(res=-1))
    m.setup(scan_1m, tgt)
    assert m.match() is True
    T = m.getResult()
    assert np.abs(T[:3, 3] - synth.T_TRUE[:3, 3]).max() < 1e-4
    assert rot_angle(T[:3, :3], synth.T_TRUE[:3, :3]) < 1e-5
    q, mm, d2 = m.correspondences()
    assert len(q) == len(scan_1m)
    assert np.mean(q == mm) > 0.9999          # every point matched to its own copy
    assert np.linalg.norm(T - synth.T_TRUE) < 0.1   # the reference tests' own bound


def test_icp_result_independent_of_point_order(W, synth):
    src, tgt = synth.scan_pair(200_000)
    rng = np.random.default_rng(0)
    ps, pt = rng.permutation(len(src)), rng.permutation(len(tgt))
    a = W.ICPMatcher(W.ICPMatcherParams(res=-1))
    a.setup(src, tgt)
    assert a.match()
    b = W.ICPMatcher(W.ICPMatcherParams(res=-1))
    b.setup(src[ps], tgt[pt])
    assert b.match()
    assert a.iterations == b.iterations
    assert np.array_equal(a.getResult(), b.getResult())    # exact sums: bit-identical transform
    qa, ma, da = a.correspondences()
    qb, mb, db = b.correspondences()
    # the same geometric pairs: source ps[qb] <-> target pt[mb]
    order = np.argsort(ps[qb])
    assert np.array_equal(ps[qb][order], qa) and np.array_equal(pt[mb][order], ma) and np.array_equal(db[order], da)


def test_point_to_plane_1m_converges_to_truth(W, synth):
    """The bench workload (config 2): two noisy 1M-point scans, analytic normals."""
    src, tgt, nrm = synth.scan_pair(1_000_000, return_normals=True)
    m = W.ICPMatcher(W.ICPMatcherParams(res=-1, estimator=W.EST_POINT_TO_PLANE))
    m.setup(src, tgt)
    m.setTargetNormals(nrm)
    assert m.match() is True
    T = m.getResult()
    # 2 cm range noise, and PCL's relative-MSE rule (fit_eps = 1e-2) stops after ~4 iterations
    assert np.abs(T[:3, 3] - synth.T_TRUE[:3, 3]).max() < 2e-2
    assert rot_angle(T[:3, :3], synth.T_TRUE[:3, :3]) < 2e-3
    assert m.iterations <= 10


def test_gicp_500k_recovers_rigid_copy(W, synth):
    """BASELINE config 3 size."""
    rings, az = synth.SIZES[500_000]
    src = synth.velodyne_scan(rings, az, None, synth.SOURCE_SEED, n_points=500_000)
    tgt = pcl_transform(src, synth.T_TRUE)
    m = W.GICPMatcher(W.GICPMatcherParams(res=-1))
    m.setup(src, tgt)
    assert m.match() is True
    T = m.getResult()
    assert np.abs(T[:3, 3] - synth.T_TRUE[:3, 3]).max() < 1e-3
    assert rot_angle(T[:3, :3], synth.T_TRUE[:3, :3]) < 1e-4


def test_ndt_1m_vs_5m_score_prefers_true_pose(W, synth, scan_1m):
    """BASELINE config 4: 1M-point scan against the 5M-point map, 0.5 m voxels."""
    from test_gpu_ndt import pose_matrix
    big = synth.map_cloud(5, 1_000_000)
    assert len(big) > 4_500_000
    m = W.NDTMatcher(W.NDTMatcherParams(res=0.5))
    m.setup(scan_1m, big)
    voxel, count, cen, mean, icov = m.grid()
    assert len(voxel) > 10_000 and (count >= 6).all() and (np.diff(voxel) > 0).all()
    Tt = synth.T_TRUE
    # pose vector of T_true in NDT's (x, y, z, Rx, Ry, Rz) convention: R = Rx Ry Rz
    ry = np.arcsin(Tt[0, 2])
    rx = np.arctan2(-Tt[1, 2], Tt[2, 2])
    rz = np.arctan2(-Tt[0, 1], Tt[0, 0])
    p_true = np.array([Tt[0, 3], Tt[1, 3], Tt[2, 3], rx, ry, rz])
    assert np.abs(pose_matrix(p_true).astype(np.float64) - Tt).max() < 1e-6
    s_true, g_true, _ = m.derivatives(p_true, pose_matrix(p_true))
    s_id, g_id, _ = m.derivatives(np.zeros(6), pose_matrix(np.zeros(6)))
    assert s_true > 1.5 * s_id
    assert np.linalg.norm(g_true[:3]) < np.linalg.norm(g_id[:3])
