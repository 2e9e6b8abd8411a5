"""The Lu-Milios information estimators: the oracle's restatement
(oracle/matcher.cpp) against golden matrices produced by executing the reference's own statements
(tests/golden/make_lum_fixture.py, from wave_matching/src/icp_pcl_functions.cpp:51-289).

Bars: information = MM * (1.0f / ss) with ss an fp32 sum that depends on D = MM^-1 MZ, and the 6x6
inverse is only specified up to rounding (Eigen's LU here, LAPACK's in the fixture, the oracle's own
elimination): MM-proportional structure to 1e-12, the overall scale 1/ss to a few fp32 ulps."""
import pathlib

import numpy as np
import pytest

from oracle import oracle as O

FIXTURE = pathlib.Path(__file__).parent / "golden" / "lum_fixture.npz"
FP32_ULPS = 4 * 2.0 ** -24


@pytest.fixture(scope="module")
def golden():
    return np.load(FIXTURE)


def _check(info, want_info, want_MM, want_ss):
    # the matrix is MM scaled by one fp32 number: its shape is pinned to double rounding ...
    scale = info[0, 0] / want_MM[0, 0]
    assert np.abs(info - want_MM * scale).max() <= 1e-12 * np.abs(info).max()
    # ... and the scale 1 / ss to fp32 resolution
    assert abs(scale * float(want_ss) - 1.0) <= FP32_ULPS
    assert np.abs(info - want_info).max() <= FP32_ULPS * np.abs(want_info).max()


@pytest.mark.parametrize("case", [0, 1, 2])
def test_lum_oracle_matches_reference_statements(golden, case):
    g = golden
    info, ok = O.estimate_lum(g[f"final{case}"], g[f"target{case}"], g[f"q{case}"], g[f"m{case}"], O.SUM_PCL)
    assert ok
    _check(info, g[f"info{case}"], g[f"MM{case}"], g[f"ss{case}"])
    # the repo's exact-sum mode agrees with the reference's sequential sums to their own rounding
    exact, ok = O.estimate_lum(g[f"final{case}"], g[f"target{case}"], g[f"q{case}"], g[f"m{case}"], O.SUM_EXACT,
                               k_quad=O.fix_scales(g[f"final{case}"], g[f"target{case}"], 3.0)[1])
    assert ok and np.abs(exact - g[f"info{case}"]).max() <= 1e-5 * np.abs(g[f"info{case}"]).max()


@pytest.mark.parametrize("case", [0, 1, 2])
def test_lum_old_oracle_matches_reference_statements(golden, case):
    g = golden
    final, target, max_corr = g[f"final{case}"], g[f"target{case}"], float(g[f"max_corr{case}"])
    if case != 1:
        assert int(g[f"old_n{case}"]) < len(final)       # the strict `d2 < max_corr^2` test drops pairs here
    info, ok = O.estimate_lum_old(final, target, max_corr, O.SUM_PCL)
    assert ok
    assert round(info[0, 0] * float(g[f"old_ss{case}"])) == int(g[f"old_n{case}"])   # MM(0,0) = numCorr
    _check(info, g[f"old_info{case}"], g[f"old_MM{case}"], g[f"old_ss{case}"])
