"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C ABI against the
CPU oracle on the same inputs.  Bars: correspondence indices and fp32 squared distances bit-exact;
per-iteration transforms, iteration count, stop rule and final transform bit-exact against the
oracle's SUM_EXACT arithmetic (the repo's estimator spec); within 1e-4 m / 1e-5 rad of
the oracle's PCL-faithful fp32 arithmetic (BASELINE.json north_star tolerance)."""
import numpy as np
import pytest

from conftest import pcl_transform, rot_angle

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("n_tgt", [1, 2, 3, 5, 9, 33])
def test_tiny_targets_duplicates_and_ties(W, oracle, n_tgt):
    """The tree degenerates for tiny targets (a one-point target has no node at all) and duplicated
    target points tie exactly: the lowest target index must win, in the stand-alone search and as the
    ICP correspondences."""
    rng = np.random.default_rng(n_tgt)
    base = rng.uniform(-2, 2, size=(n_tgt, 3)).astype(np.float32)
    tgt = np.concatenate([base, base[::-1]])             # every point twice: exact ties everywhere
    qry = rng.uniform(-3, 3, size=(257, 3)).astype(np.float32)
    qry[:n_tgt] = base                                    # zero-distance queries too
    idx, d2 = W.NearestNeighbour(tgt).search(qry)
    ridx, rd2 = oracle.brute_nn1(tgt, qry)
    assert np.array_equal(idx, ridx) and np.array_equal(d2, rd2)
    assert (idx < n_tgt).all()                            # never the duplicate with the higher index
    if n_tgt >= 3:
        m = W.ICPMatcher(W.ICPMatcherParams(res=-1, max_corr=100.0, max_iter=3))
        m.setup(qry, tgt)
        m.match()
        ref = oracle.icp_align(qry, tgt, max_corr=100.0, max_iter=3, sum_mode=oracle.SUM_EXACT)
        q, mm, dd = m.correspondences()
        assert m.iterations == ref.iterations
        assert np.array_equal(mm, ref.corr_match) and np.array_equal(dd, ref.corr_dist)


@pytest.mark.parametrize("shape", ["slab", "line", "clusters", "far_queries"])
def test_entry_table_adversarial_geometry(W, oracle, shape):
    """The entry table (coarse Morton cells -> subtrees) and the Morton-key seeds must not change a single
    correspondence: degenerate extents, clustered duplicates, queries far outside the target box and
    bounds that span many cells, checked bit for bit against the oracle over several iterations."""
    rng = np.random.default_rng({"slab": 1, "line": 2, "clusters": 3, "far_queries": 4}[shape])
    n = 6000
    if shape == "slab":          # 200 m x 150 m x 2 cm
        tgt = rng.uniform([-100, -75, -0.01], [100, 75, 0.01], size=(n, 3))
    elif shape == "line":        # zero extent on two axes
        tgt = np.zeros((n, 3))
        tgt[:, 0] = rng.uniform(-50, 50, n)
    elif shape == "clusters":    # a few tight clusters plus exact duplicates
        centers = rng.uniform(-60, 60, size=(12, 3))
        tgt = centers[rng.integers(0, 12, n)] + rng.normal(0, 0.02, (n, 3))
        tgt[n // 2:] = tgt[: n - n // 2]
    else:
        tgt = rng.uniform(-5, 5, size=(n, 3))
    tgt = tgt.astype(np.float32)
    src = (tgt[rng.permutation(n)] + rng.normal(0, 0.3, (n, 3))).astype(np.float32)
    if shape == "far_queries":
        src[::7] += np.float32(40.0)             # far outside the target's bounding box
    src[::11] += rng.uniform(-2.5, 2.5, (len(src[::11]), 3)).astype(np.float32)
    kw = dict(max_corr=5.0 if shape != "far_queries" else 100.0, max_iter=5)
    m = W.ICPMatcher(W.ICPMatcherParams(res=-1, **kw))
    m.setup(src, tgt)
    m.match()
    ref = oracle.icp_align(src, tgt, sum_mode=oracle.SUM_EXACT, nn_threads=8, **kw)
    q, mm, dd = m.correspondences()
    assert m.iterations == ref.iterations
    assert np.array_equal(q, ref.corr_query) and np.array_equal(mm, ref.corr_match) and np.array_equal(dd, ref.corr_dist)
    if ref.converged:
        assert np.array_equal(m.getResult().astype(np.float32), ref.T)


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


# ---- voxel grid, multiscale and information matrices (reference tests restated) --------------------
@pytest.mark.parametrize("leaf", [0.05, 0.1, 0.8])
def test_voxel_grid_bit_exact(W, oracle, testscan, leaf):
    got, ok = W.voxel_grid(testscan, leaf)
    ref, rok = oracle.voxel_grid(testscan, leaf)
    assert ok == rok and got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_voxel_grid_overflow_passthrough(W, oracle, testscan):
    got, ok = W.voxel_grid(testscan[:5000], 1e-4)  # dx*dy*dz > INT_MAX: PCL returns the input
    ref, rok = oracle.voxel_grid(testscan[:5000], 1e-4)
    assert ok is False and rok is False
    assert np.array_equal(got, ref)


CASES = {
    # name: (res, multiscale_steps, tx)   tests/icp_tests.cpp:65-148 with tests/config/icp.yaml
    "nullDisplacement": (0.05, 0, 0.0),
    "smallDisplacement": (0.05, 0, 0.2),
    "multiscale": (0.1, 3, 0.2),
}


@pytest.mark.parametrize("name", list(CASES))
def test_reference_icp_cases_bit_exact(W, oracle, testscan, name):
    res, steps, tx = CASES[name]
    T = np.eye(4)
    T[0, 3] = tx
    tgt = pcl_transform(testscan, T)
    m = W.ICPMatcher(W.ICPMatcherParams(res=res, multiscale_steps=steps))
    m.setup(testscan, tgt)
    assert m.match() is True
    assert np.linalg.norm(m.getResult() - T) < 0.1  # tests/icp_tests.cpp:37
    ref = oracle.icp_match(testscan, tgt, res=res, multiscale_steps=steps, nn_threads=8)
    assert ref.success
    assert m.iterations == ref.total_iterations
    assert np.array_equal(m.getResult(), ref.T)
    q, mm, d2 = m.correspondences()
    assert np.array_equal(q, ref.last.corr_query)
    assert np.array_equal(mm, ref.last.corr_match)
    assert np.array_equal(d2, ref.last.corr_dist)
    assert np.array_equal(m.aligned()[:, :3], ref.last.aligned[:, :3])


def test_information_lum_vs_lumold(W, oracle, testscan):
    """smallinfo + lumvslum (tests/icp_tests.cpp:105-195): one scan is distorted by U(-0.3, 0.3)."""
    rng = np.random.default_rng(0)
    T = np.eye(4)
    T[0, 3] = 0.2
    tgt = pcl_transform(testscan, T).astype(np.float64)
    tgt = (tgt + rng.uniform(-0.3, 0.3, tgt.shape)).astype(np.float32)
    m = W.ICPMatcher(W.ICPMatcherParams(res=0.05, multiscale_steps=0, covar_estimator=W.INFO_LUMOLD))
    m.setup(testscan, tgt)
    m.match()
    lum, lumold = m.info(W.INFO_LUM), m.info(W.INFO_LUMOLD)
    assert lumold[0, 0] > 0
    ref = oracle.icp_match(testscan, tgt, res=0.05, multiscale_steps=0, nn_threads=8)
    k_quad = oracle.fix_scales(ref.ds_ref, ref.ds_tgt, 3.0)[1]
    r_lum, _ = oracle.estimate_lum(ref.last.aligned, ref.ds_tgt, ref.last.corr_query, ref.last.corr_match,
                                   oracle.SUM_EXACT, k_quad)
    r_old, _ = oracle.estimate_lum_old(ref.last.aligned, ref.ds_tgt, 3.0, oracle.SUM_EXACT, k_quad, nn_threads=8)
    assert np.array_equal(lum, r_lum)
    assert np.array_equal(lumold, r_old)
    # against the reference's own (sequential fp32/fp64) arithmetic
    p_old, _ = oracle.estimate_lum_old(ref.last.aligned, ref.ds_tgt, 3.0, oracle.SUM_PCL, nn_threads=8)
    assert np.allclose(lumold, p_old, rtol=1e-4)
    # estimateInfo(): the switch falls through, the result is always LUMold's (src/icp.cpp:135-142)
    m.params.covar_estimator = W.INFO_LUM
    m.estimateInfo()
    assert np.array_equal(m.getInfo(), lumold)


@pytest.mark.parametrize("mode", ["fullres", "voxel", "multiscale"])
def test_information_censi_vs_oracle(W, oracle, testscan, mode):
    """estimateCensi (src/icp.cpp:167-397) over the correspondences, clouds and result of the last
    match, in the three branches of match().  fp64 sums in a different order and the device's
    atan2f / atanf (an ulp off glibc's now and then) bound the agreement, hence 1e-6 relative."""
    T = np.eye(4)
    T[:3, :3] = rot_z(np.deg2rad(-1.5)) @ rot_x(np.deg2rad(-0.8))     # negative roll: folded Euler angles
    T[:3, 3] = (0.2, -0.1, 0.05)
    tgt = pcl_transform(testscan, T)
    kw = {"fullres": dict(res=-1), "voxel": dict(res=0.1, multiscale_steps=0),
          "multiscale": dict(res=0.1, multiscale_steps=2)}[mode]
    m = W.ICPMatcher(W.ICPMatcherParams(covar_estimator=W.INFO_CENSI, **kw))
    m.setup(testscan, tgt)
    assert m.match()
    got = m.info(W.INFO_CENSI)
    if mode == "fullres":
        ref_cloud, tgt_cloud = testscan, tgt
        q, mm, _ = m.correspondences()
    else:
        ref = oracle.icp_match(testscan, tgt, nn_threads=8, **kw)
        ref_cloud, tgt_cloud, q, mm = ref.ds_ref, ref.ds_tgt, ref.last.corr_query, ref.last.corr_match
        gq, gm, _ = m.correspondences()
        assert np.array_equal(gq, q) and np.array_equal(gm, mm)
    want, H, M, ok = oracle.estimate_censi(ref_cloud, tgt_cloud, q, mm, m.getResult())
    assert ok and np.all(np.isfinite(got))
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    assert want[0, 0] > 0


def rot_x(a):
    return np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])


def rot_z(a):
    return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])


def test_point_to_plane_bit_exact_and_estimated_normals(W, oracle):
    """SURVEY.md 8(a) A7 (north-star extension; the reference has no point-to-plane estimator):
    with supplied normals the GPU follows the oracle's PointToPlaneLLS restatement bit for bit; with
    no normals it estimates them on the device (k-NN PCA) and must land on the oracle's result for
    the oracle's own estimate of the same normals."""
    from libwave_b200 import synth
    src, tgt, nrm = synth.scan_pair(10_000, return_normals=True)
    m = W.ICPMatcher(W.ICPMatcherParams(res=-1, estimator=W.EST_POINT_TO_PLANE))
    m.setup(src, tgt)
    m.setTargetNormals(nrm)
    assert m.match()
    ref = oracle.icp_align(src, tgt, estimator=oracle.EST_POINT_TO_PLANE, sum_mode=oracle.SUM_EXACT, target_normals=nrm)
    mse, ncorr, Ttr = m.trace()
    assert m.iterations == ref.iterations and np.array_equal(ncorr, ref.n_corr)
    assert np.array_equal(mse, ref.mse)
    assert np.abs(Ttr - ref.T_trace).max() < 1e-7      # cos/sin of the device vs glibc: <= 1 fp32 ulp
    assert np.abs(m.getResult() - ref.T.astype(np.float64)).max() < 1e-6
    q, mm, d2 = m.correspondences()
    assert np.array_equal(mm, ref.corr_match)
    pcl = oracle.icp_align(src, tgt, estimator=oracle.EST_POINT_TO_PLANE, sum_mode=oracle.SUM_PCL, target_normals=nrm)
    assert np.abs(m.getResult()[:3, 3] - pcl.T[:3, 3]).max() < 1e-4
    assert rot_angle(m.getResult()[:3, :3], pcl.T[:3, :3]) < 1e-5

    m2 = W.ICPMatcher(W.ICPMatcherParams(res=-1, estimator=W.EST_POINT_TO_PLANE))
    m2.setup(src, tgt)                                   # no normals given
    assert m2.match()
    est = oracle.estimate_normals(tgt, 10)
    ref2 = oracle.icp_align(src, tgt, estimator=oracle.EST_POINT_TO_PLANE, sum_mode=oracle.SUM_EXACT,
                            target_normals=est)
    assert m2.iterations == ref2.iterations
    assert np.abs(m2.getResult()[:3, 3] - ref2.T[:3, 3]).max() < 1e-4
    assert rot_angle(m2.getResult()[:3, :3], ref2.T[:3, :3]) < 1e-5
