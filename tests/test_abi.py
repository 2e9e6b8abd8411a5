"""CPU tests (-m "not gpu"): the C-ABI library builds, loads and exports every symbol that
include/wavecu.h declares; the ctypes table matches the header; no compute calls (no GPU here)."""
import ctypes
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def built():
    from libwave_b200 import build
    return build.build()


def declared_functions():
    text = (ROOT / "include" / "wavecu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wavecu_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_groups():
    names = declared_functions()
    for must in ("wavecu_icp_create", "wavecu_icp_set_source", "wavecu_icp_set_target", "wavecu_icp_match",
                 "wavecu_icp_align", "wavecu_icp_correspondences", "wavecu_icp_info", "wavecu_nn_search",
                 "wavecu_voxel_grid", "wavecu_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(str(built))
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in wavecu.h but not exported: {missing}"


def test_ctypes_table_covers_the_header(built):
    from libwave_b200 import capi
    assert sorted(capi.SIGNATURES) == declared_functions()
    capi.lib()  # loads and binds every signature


def test_params_struct_layout_matches_header(built):
    from libwave_b200 import capi
    lib = capi.lib()
    p = capi.IcpParamsC()
    lib.wavecu_icp_default_params(ctypes.byref(p))
    # wave::ICPMatcherParams defaults, icp.hpp:35-64
    assert (p.max_corr, p.max_iter, p.t_eps, p.fit_eps) == (3.0, 100, 1e-8, 1e-2)
    assert (p.lidar_ang_covar, p.lidar_lin_covar, p.multiscale_steps) == (7.78e-9, 2.5e-4, 3)
    assert abs(p.res - 0.1) < 1e-7 and p.covar_estimator == 0 and p.estimator == 0


def test_no_cpu_fallback_without_a_device(built):
    """On a box without a GPU every entry point must fail loudly instead of computing on the CPU."""
    from libwave_b200 import capi
    lib = capi.lib()
    if lib.wavecu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = ctypes.c_void_p()
    assert lib.wavecu_icp_create(None, 0, None, ctypes.byref(h)) == -2  # WAVECU_ERR_CUDA
    assert b"no CPU fallback" in lib.wavecu_last_error()
    import libwave_b200 as W
    with pytest.raises(capi.WavecuError):
        W.ICPMatcher()
    with pytest.raises(capi.WavecuError):
        W.NearestNeighbour()


def test_product_never_imports_the_oracle():
    for path in list((ROOT / "libwave_b200").rglob("*.py")) + list((ROOT / "libwave_b200" / "csrc").glob("*")) + \
            list((ROOT / "include").rglob("*")) + list((ROOT / "src").rglob("*")):
        if path.is_file() and path.suffix in (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp"):
            text = path.read_text(errors="ignore")
            bad = re.search(r"(^|\n)\s*(from\s+oracle|import\s+oracle)|#include\s*[<\"][^>\"]*oracle|libwave_oracle|wo_[a-z]+_",
                            text)
            assert not bad, f"{path} reaches into the oracle: {bad.group(0)!r}"


def test_cpp_shim_builds_and_links(built):
    """The C++ drop-in (libwave_matching.so) and the native test binary build with g++ here."""
    from libwave_b200 import build
    so = build.build_shim()
    assert so.exists() and (ROOT / "tests" / "cpp" / "_build" / "test_matching").exists()
    lib = ctypes.CDLL(str(built), mode=ctypes.RTLD_GLOBAL)
    assert lib is not None
    shim = ctypes.CDLL(str(so))
    # the mangled constructor / match of wave::ICPMatcher must be exported
    import subprocess
    syms = subprocess.run(["nm", "-DC", str(so)], capture_output=True, text=True).stdout
    for name in ("wave::ICPMatcher::match()", "wave::ICPMatcher::setRef", "wave::ICPMatcher::estimateInfo()",
                 "wave::ICPMatcherParams::ICPMatcherParams(", "wave::ConfigParser::load("):
        assert name in syms, name
    assert shim is not None
