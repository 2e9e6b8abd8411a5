"""The Censi information estimator: the oracle's restatement against golden
matrices produced by executing the reference's own statements (tests/golden/make_censi_fixture.py)."""
import pathlib

import numpy as np
import pytest

from oracle import oracle as O

FIXTURE = pathlib.Path(__file__).parent / "golden" / "censi_fixture.npz"


@pytest.fixture(scope="module")
def golden():
    return np.load(FIXTURE)


@pytest.mark.parametrize("case", [0, 1, 2])
def test_censi_oracle_matches_reference_statements(golden, case):
    g = golden
    info, H, M, ok = O.estimate_censi(g[f"ref{case}"], g[f"target{case}"], g[f"q{case}"], g[f"m{case}"], g[f"T{case}"],
                                      float(g["lin"]), float(g["ang"]))
    assert ok
    # The expanded closed forms of the reference and the assembled derivatives agree to rounding.  The
    # largest gap (~2e-9 of the matrix scale) sits in d2J_dX2(3,5) and (5,5), whose reference
    # expressions contain fp32 products of two coordinates (`2 * Z3 * Z4 * ...` with float Z's) that the
    # assembled form evaluates in double.
    for got, want, name in ((H, g[f"H{case}"], "d2J_dX2"), (M, g[f"middle{case}"], "middle")):
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= 1e-7 * scale, name
    # everywhere else d2J_dX2 agrees to double rounding; `middle` carries the fp32 angles atan2f / atanf,
    # where numpy (the fixture) and glibc (the oracle) differ by an ulp now and then
    want_H, scale = g[f"H{case}"], np.abs(g[f"H{case}"]).max()
    off = [(i, j) for i in range(6) for j in range(6) if (i, j) not in ((3, 5), (5, 3), (5, 5))]
    assert max(abs(H[i, j] - want_H[i, j]) for i, j in off) <= 1e-11 * scale
    want = g[f"info{case}"]
    assert np.abs(info - want).max() <= 1e-6 * np.abs(want).max()
    assert np.allclose(info, info.T, rtol=0, atol=1e-6 * np.abs(want).max())


def test_censi_euler_fold(golden):
    """case 1 has a negative rotation about x: Eigen's eulerAngles(0,1,2) folds the first angle into
    [0, pi], and the estimator must be evaluated at those folded angles."""
    e = golden["euler1"]
    assert 0.0 <= e[0] <= np.pi and e[0] > 3.0
