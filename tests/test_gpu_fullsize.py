"""GPU tests at BASELINE.json's full sizes: the ICP configurations against the oracle itself (with
all host threads its exact NN search takes seconds at 1 M points), and size-independent properties:
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


def _host_threads():
    import os
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def test_bench_workload_1m_point_to_plane_equals_oracle(W, oracle, synth):
    """BASELINE config 2 at its own size - the exact workload bench.py times: iteration count, MSE
    trace, correspondence counts, every correspondence index and fp32 distance equal to the oracle's
    (integer / fp32 work: bit-exact); transforms within one fp32 ulp of the oracle's exact-sum mode
    (device cos/sin vs glibc) and within the north-star tolerance (1e-4 m, 1e-5 rad) of its
    PCL-faithful summation mode."""
    src, tgt, nrm = synth.scan_pair(1_000_000, return_normals=True)
    m = W.ICPMatcher(W.ICPMatcherParams(res=-1, estimator=W.EST_POINT_TO_PLANE))
    m.setup(src, tgt)
    m.setTargetNormals(nrm)
    assert m.match() is True
    nt = _host_threads()
    ref = oracle.icp_align(src, tgt, estimator=oracle.EST_POINT_TO_PLANE, sum_mode=oracle.SUM_EXACT,
                           target_normals=nrm, nn_threads=nt)
    mse, ncorr, Ttr = m.trace()
    assert ref.converged and m.iterations == ref.iterations
    assert np.array_equal(ncorr, ref.n_corr)
    assert np.array_equal(mse, ref.mse)
    assert np.abs(Ttr - ref.T_trace).max() < 1e-7
    assert np.abs(m.getResult() - ref.T.astype(np.float64)).max() < 1e-6
    q, mm, d2 = m.correspondences()
    assert np.array_equal(q, ref.corr_query)
    assert np.array_equal(mm, ref.corr_match)
    assert np.array_equal(d2, ref.corr_dist)
    pcl = oracle.icp_align(src, tgt, estimator=oracle.EST_POINT_TO_PLANE, sum_mode=oracle.SUM_PCL,
                           target_normals=nrm, nn_threads=nt)
    assert m.iterations == pcl.iterations
    assert np.abs(m.getResult()[:3, 3] - pcl.T[:3, 3]).max() < 1e-4
    assert rot_angle(m.getResult()[:3, :3], pcl.T[:3, :3]) < 1e-5


def test_icp_1m_svd_equals_oracle(W, oracle, synth):
    """The reference's own estimator (SVD, icp.hpp:100) on the 1 M / 1 M pair: everything bit-exact
    against the oracle's exact-sum mode, aligned cloud included."""
    src, tgt = synth.scan_pair(1_000_000)
    m = W.ICPMatcher(W.ICPMatcherParams(res=-1))
    m.setup(src, tgt)
    assert m.match() is True
    nt = _host_threads()
    ref = oracle.icp_align(src, tgt, sum_mode=oracle.SUM_EXACT, nn_threads=nt)
    mse, ncorr, Ttr = m.trace()
    assert m.iterations == ref.iterations
    assert np.array_equal(ncorr, ref.n_corr) and np.array_equal(mse, ref.mse) and np.array_equal(Ttr, ref.T_trace)
    assert np.array_equal(m.getResult().astype(np.float32), ref.T)
    q, mm, d2 = m.correspondences()
    assert np.array_equal(q, ref.corr_query) and np.array_equal(mm, ref.corr_match) and np.array_equal(d2, ref.corr_dist)
    assert np.array_equal(m.aligned()[:, :3], ref.aligned[:, :3])
    # Against the PCL-faithful summation mode the bar at this size is the reference estimator's own
    # noise: Umeyama's fp32 sums over 10^6 pairs (restated as strictly sequential) move
    # each incremental transform by millimetres (measured: 3.9 mm after the first iteration, 3.3 mm at
    # the end, same 9 iterations) - the 1e-4 m / 1e-5 rad bar is met where PCL itself sums in fp64
    # (point-to-plane, the test above) and on the reference's 55 k-point fixture.
    pcl = oracle.icp_align(src, tgt, sum_mode=oracle.SUM_PCL, nn_threads=nt)
    assert m.iterations == pcl.iterations
    assert np.abs(m.getResult()[:3, 3] - pcl.T[:3, 3]).max() < 1e-2
    assert rot_angle(m.getResult()[:3, :3], pcl.T[:3, :3]) < 1e-3


@pytest.mark.parametrize("scan_id", [0, 255])
def test_batch_scan_200k_equals_oracle(W, oracle, synth, scan_id):
    """BASELINE config 5: one 200 k-point scan-to-map alignment of the batch (first and last scan),
    SVD estimator, bit-exact against the oracle."""
    sources, target = synth.scan_batch(200_000, 0, ids=[scan_id])
    src = sources[0]
    m = W.ICPMatcher(W.ICPMatcherParams(res=-1))
    m.setup(src, target)
    ok = m.match()
    ref = oracle.icp_align(src, target, sum_mode=oracle.SUM_EXACT, nn_threads=_host_threads())
    assert ok == ref.converged and m.iterations == ref.iterations
    mse, ncorr, Ttr = m.trace()
    assert np.array_equal(ncorr, ref.n_corr) and np.array_equal(mse, ref.mse) and np.array_equal(Ttr, ref.T_trace)
    assert np.array_equal(m.getResult().astype(np.float32), ref.T)
    q, mm, d2 = m.correspondences()
    assert np.array_equal(q, ref.corr_query) and np.array_equal(mm, ref.corr_match) and np.array_equal(d2, ref.corr_dist)


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


def test_gicp_500k_noisy_pair_equals_oracle(W, oracle, synth):
    """BASELINE config 3 at its own size and on the bench's own inputs (two noisy 500 k-point scans):
    the GPU follows the oracle's BFGS trajectory exactly - outer iterations, cost evaluations,
    correspondences and the final transform, bit for bit (exact sums on both sides; the oracle needs
    ~25 s for this case).  How close that transform is to the truth is PCL GICP's own property: with
    k = 10 neighbours on a 7813-step ring every neighbourhood is a line segment, the plane-to-plane
    model degenerates, and the restated algorithm stops 16 cm from the truth (1.6 cm / 0.6 cm at
    10 k / 200 k points) - recorded here, not asserted away."""
    src, tgt = synth.scan_pair(500_000)
    m = W.GICPMatcher(W.GICPMatcherParams(res=-1))
    m.setup(src, tgt)
    ok = m.match()
    ref = oracle.gicp_align(src, tgt)
    st = m.stats()
    assert ok == ref.converged and m.iterations == ref.iterations
    assert st["evaluations"] == ref.evaluations and st["n_corr"] == ref.n_corr
    assert np.array_equal(m.getResult().astype(np.float32), ref.T)
    assert np.abs(m.getResult()[:3, 3] - synth.T_TRUE[:3, 3]).max() < 0.25


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
