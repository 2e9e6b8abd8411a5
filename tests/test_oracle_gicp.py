"""CPU tests (-m "not gpu") of the GICP oracle (oracle/gicp.cpp): the reference's test cases re-stated
(tests/gicp_tests.cpp:43-100, bound 0.1) and the structure of the per-point covariances."""
import numpy as np

from conftest import pcl_transform


def test_gicp_covariances_encode_surface_normals(oracle):
    rng = np.random.default_rng(0)
    pts = np.zeros((400, 3), np.float32)                     # a plane z = 0.3 x with small noise
    pts[:, :2] = rng.uniform(-1, 1, (400, 2))
    pts[:, 2] = 0.3 * pts[:, 0] + rng.normal(0, 1e-3, 400)
    covs = oracle.gicp_covariances(pts, 10, 1e-3)
    ev, V = np.linalg.eigh(covs)
    assert np.allclose(ev[:, 0], 1e-3, atol=1e-9) and np.allclose(ev[:, 1:], 1.0, atol=1e-9)
    n = np.array([-0.3, 0.0, 1.0]) / np.hypot(0.3, 1.0)
    assert np.median(np.abs(V[:, :, 0] @ n)) > 0.999         # smallest direction = plane normal


def test_gicp_reference_cases_restated(oracle, testscan):
    for res, tx, iters in ((-1.0, 0.0, 1), (0.05, 0.0, 1), (0.05, 0.2, None)):
        T = np.eye(4)
        T[0, 3] = tx
        tgt = pcl_transform(testscan, T)
        s, t = (testscan, tgt) if res <= 0 else (oracle.voxel_grid(testscan, res)[0], oracle.voxel_grid(tgt, res)[0])
        r = oracle.gicp_align(s, t)
        assert r.converged
        assert np.linalg.norm(r.T - T) < 0.1                 # tests/gicp_tests.cpp:59,79,99
        assert np.linalg.norm(r.T - T) < 1e-3
        if iters is not None:
            assert r.iterations == iters
    assert not oracle.gicp_align(np.zeros((0, 3), np.float32), testscan).converged
    assert not oracle.gicp_align(testscan[:3], testscan).converged   # < 4 pairs: PCL throws inside align
