"""GPU parity tests for NDTMatcher through the C ABI against the oracle.

NDT is floating point end to end (exp, fp64 sums, Newton steps), so the bars are tolerances,
written here: voxel membership and counts exact; cell statistics 1e-9 relative; score / gradient /
Hessian of one derivative pass 1e-9 relative to their largest entry; final transforms within the
north-star tolerance (1e-4 m, 1e-5 rad) where the optimisation is well conditioned."""
import numpy as np
import pytest

from conftest import pcl_transform, rot_angle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def W():
    import libwave_b200 as W
    return W


def pose_matrix(p):
    """(Translation * AngleAxis(rx,X) * AngleAxis(ry,Y) * AngleAxis(rz,Z)).matrix() in fp32."""
    rx, ry, rz = (np.float32(v) for v in p[3:])
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    T = np.eye(4, dtype=np.float32)
    T[0, :3] = [cy * cz, -cy * sz, sy]
    T[1, :3] = [sx * sy * cz + cx * sz, -sx * sy * sz + cx * cz, -sx * cy]
    T[2, :3] = [-cx * sy * cz + sx * sz, cx * sy * sz + sx * cz, cx * cy]
    T[:3, 3] = np.asarray(p[:3], dtype=np.float32)
    return T


@pytest.mark.parametrize("res", [0.3, 1.0])
def test_ndt_grid_matches_oracle(W, oracle, testscan, res):
    m = W.NDTMatcher(W.NDTMatcherParams(res=res))
    m.setTarget(testscan)
    voxel, count, cen, mean, icov = m.grid()
    rv, rc, rcen, rmean, ricov = oracle.ndt_grid(testscan, res)
    assert np.array_equal(voxel, rv) and np.array_equal(count, rc)
    assert np.array_equal(cen, rcen)                       # fp32 centroid: same summation order
    assert np.allclose(mean, rmean, rtol=1e-12, atol=1e-12)
    scale = np.abs(ricov).max(axis=(1, 2), keepdims=True)
    assert (np.abs(icov - ricov) <= 1e-9 * scale).all()


@pytest.mark.parametrize("res,pose", [(0.3, [0, 0, 0, 0, 0, 0]), (1.0, [0.05, 0.02, -0.01, 0.01, -0.02, 0.015]),
                                      (0.5, [-0.1, 0.03, 0.02, 0.0, 0.005, -0.03])])
def test_ndt_derivatives_match_oracle(W, oracle, testscan, res, pose):
    T0 = np.eye(4)
    T0[0, 3] = 0.2
    tgt = pcl_transform(testscan, T0)
    m = W.NDTMatcher(W.NDTMatcherParams(res=res))
    m.setup(testscan, tgt)
    T = pose_matrix(pose)
    s, g, H = m.derivatives(pose, T)
    rs, rg, rH = oracle.ndt_derivatives(testscan, tgt, res, pose, T)
    assert abs(s - rs) <= 1e-9 * abs(rs)
    assert np.abs(g - rg).max() <= 1e-9 * np.abs(rg).max()
    assert np.abs(H - rH).max() <= 1e-9 * np.abs(rH).max()


@pytest.mark.parametrize("res", [0.05, 0.1])
def test_ndt_reference_null_cases(W, oracle, testscan, res):
    """fullResNullMatch (res from tests/config/ndt.yaml = 0.05) and nullDisplacement (res = 0.1),
    tests/ndt_tests.cpp:45-82: match()==true and ||result - I||_F < 0.12."""
    m = W.NDTMatcher(W.NDTMatcherParams(res=res))
    m.setup(testscan, testscan.copy())
    assert m.match() is True
    assert np.linalg.norm(m.getResult() - np.eye(4)) < 0.12
    ref = oracle.ndt_align(testscan, testscan.copy(), res=res)
    assert ref.converged and m.iterations == ref.iterations
    assert np.abs(m.getResult()[:3, 3] - ref.T[:3, 3]).max() < 1e-4
    assert rot_angle(m.getResult()[:3, :3], ref.T[:3, :3]) < 1e-5


def test_ndt_reference_small_displacement(W, oracle, testscan):
    """smallDisplacement (tests/ndt_tests.cpp:85-102): res 0.3, target moved by 0.2 m in x, match()
    true and ||result - perturb||_F < 0.12 - with the line search of PCL >= 1.9 (the default, see
    wavecu.h); iteration count equal to the oracle's, transform within the north-star tolerance."""
    T0 = np.eye(4)
    T0[0, 3] = 0.2
    tgt = pcl_transform(testscan, T0)
    m = W.NDTMatcher(W.NDTMatcherParams(res=0.3))
    m.setup(testscan, tgt)
    assert m.match() is True
    assert np.linalg.norm(m.getResult() - T0) < 0.12
    ref = oracle.ndt_align(testscan, tgt, res=0.3)
    assert ref.converged and m.iterations == ref.iterations
    assert np.abs(m.getResult()[:3, 3] - ref.T[:3, 3]).max() < 1e-4
    assert rot_angle(m.getResult()[:3, :3], ref.T[:3, :3]) < 1e-5


def test_ndt_pcl18_first_steps_follow_oracle(W, oracle, testscan):
    """The same inputs with PCL 1.8's skipped line search: a capped Newton iteration on indefinite
    Hessians, chaotic over its 102 iterations, so the comparison is made where it is meaningful -
    after a few steps."""
    T0 = np.eye(4)
    T0[0, 3] = 0.2
    tgt = pcl_transform(testscan, T0)
    for iters in (1, 2, 3):
        m = W.NDTMatcher(W.NDTMatcherParams(res=0.3, max_iter=iters - 2, line_search=W.NDT_LS_PCL18))
        m.setup(testscan, tgt)
        m.match()
        ref = oracle.ndt_align(testscan, tgt, res=0.3, max_iter=iters - 2, line_search=oracle.NDT_LS_PCL18)
        assert m.iterations == ref.iterations == iters
        assert np.abs(m.getResult() - ref.T).max() < 1e-5


def test_ndt_stats_counts_cells(W, oracle, testscan):
    m = W.NDTMatcher(W.NDTMatcherParams(res=0.5))
    m.setup(testscan, testscan.copy())
    voxel = m.grid()[0]
    assert m.stats()["n_cells"] == len(voxel) == len(oracle.ndt_grid(testscan, 0.5)[0])


def test_ndt_degenerate_inputs(W, testscan):
    m = W.NDTMatcher(W.NDTMatcherParams(res=0.01))   # clamped to min_res = 0.05 (src/ndt.cpp:23-26)
    assert abs(m.getRes() - 0.05) < 1e-9
    m2 = W.NDTMatcher(W.NDTMatcherParams(res=1.0))
    m2.setup(np.zeros((0, 3), np.float32), testscan)
    assert m2.match() is False
    m2.setup(testscan, np.zeros((0, 3), np.float32))
    assert m2.match() is False
