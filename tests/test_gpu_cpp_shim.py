"""GPU test: the reference's own gtest cases (tests/icp_tests.cpp, tests/multi_matcher_tests.cpp)
re-stated in C++ against the drop-in headers (include/wave/matching/*.hpp) and run as a native
binary - the path a libwave user would take (no Python in it)."""
import pathlib
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def test_reference_gtest_cases_native():
    from libwave_b200 import build
    build.build()
    exe = ROOT / "tests" / "cpp" / "_build" / "test_matching"
    assert exe.exists()
    r = subprocess.run([str(exe), str(ROOT / "tests" / "golden" / "testscan_xyz.f32"),
                        str(ROOT / "tests" / "cpp" / "config" / "icp.yaml"),
                        str(ROOT / "tests" / "cpp" / "config" / "ndt.yaml"),
                        str(ROOT / "tests" / "cpp" / "config" / "gicp.yaml")], capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "PASSED" in r.stdout
