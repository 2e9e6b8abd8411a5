"""ctypes declarations of include/wavecu.h.  Loading fails loudly when libwavecu.so is missing:
there is no CPU or PyTorch fallback for this path."""
from __future__ import annotations

import ctypes as C
import pathlib

_DIR = pathlib.Path(__file__).resolve().parent
import os

SO = _DIR / os.environ.get("WAVECU_SO", "libwavecu.so")  # WAVECU_SO: tuning builds only

EST_SVD, EST_POINT_TO_PLANE = 0, 1
INFO_LUM, INFO_CENSI, INFO_LUMOLD = 0, 1, 2
CONV_STATES = ("NOT_CONVERGED", "ITERATIONS", "TRANSFORM", "ABS_MSE", "REL_MSE", "NO_CORRESPONDENCES")


class WavecuError(RuntimeError):
    pass


class IcpParamsC(C.Structure):
    """wavecu_icp_params == wave::ICPMatcherParams (icp.hpp:30-65) + estimator."""
    _fields_ = [("max_corr", C.c_double), ("max_iter", C.c_int), ("t_eps", C.c_double), ("fit_eps", C.c_double),
                ("lidar_ang_covar", C.c_double), ("lidar_lin_covar", C.c_double), ("multiscale_steps", C.c_int),
                ("res", C.c_float), ("covar_estimator", C.c_int), ("estimator", C.c_int)]


class NdtParamsC(C.Structure):
    """wavecu_ndt_params == wave::NDTMatcherParams (ndt.hpp:37-41) + the PCL line-search switch."""
    _fields_ = [("step_size", C.c_int), ("max_iter", C.c_int), ("t_eps", C.c_double), ("res", C.c_float),
                ("line_search", C.c_int)]


class GicpParamsC(C.Structure):
    """wavecu_gicp_params == wave::GICPMatcherParams (gicp.hpp:34-38)."""
    _fields_ = [("corr_rand", C.c_int), ("max_iter", C.c_int), ("r_eps", C.c_double), ("fit_eps", C.c_double),
                ("res", C.c_float)]


class BatchRecordC(C.Structure):
    """wavecu_batch_record: one scan's result (Matcher::result, Matcher::information, flags)."""
    _fields_ = [("T", C.c_double * 16), ("info", C.c_double * 36), ("converged", C.c_int), ("iterations", C.c_int),
                ("scan_id", C.c_int), ("device", C.c_int)]


class StatsC(C.Structure):
    _fields_ = [("build_ms", C.c_double), ("iterate_ms", C.c_double), ("solve_ms", C.c_double),
                ("total_ms", C.c_double), ("iterate_launches", C.c_longlong), ("kernel_launches", C.c_longlong),
                ("pairs", C.c_longlong), ("fallback_queries", C.c_longlong)]


_fp, _ip, _dp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)
_vp, _sz = C.c_void_p, C.c_size_t
_szp = C.POINTER(C.c_size_t)

# name -> (restype, argtypes); tests/test_abi.py checks this table against include/wavecu.h
SIGNATURES = {
    "wavecu_icp_default_params": (None, [C.POINTER(IcpParamsC)]),
    "wavecu_icp_create": (C.c_int, [C.POINTER(IcpParamsC), C.c_int, _vp, C.POINTER(_vp)]),
    "wavecu_icp_destroy": (C.c_int, [_vp]),
    "wavecu_icp_set_params": (C.c_int, [_vp, C.POINTER(IcpParamsC)]),
    "wavecu_icp_set_source": (C.c_int, [_vp, _fp, _sz]),
    "wavecu_icp_set_target": (C.c_int, [_vp, _fp, _sz]),
    "wavecu_icp_set_target_normals": (C.c_int, [_vp, _fp, _sz]),
    "wavecu_icp_set_source_device": (C.c_int, [_vp, _vp, _sz]),
    "wavecu_icp_set_target_device": (C.c_int, [_vp, _vp, _sz]),
    "wavecu_icp_set_target_normals_device": (C.c_int, [_vp, _vp, _sz]),
    "wavecu_icp_align": (C.c_int, [_vp, _dp, _ip, _ip, _ip]),
    "wavecu_icp_match": (C.c_int, [_vp, _dp, _ip, _ip]),
    "wavecu_icp_correspondences": (C.c_int, [_vp, _ip, _ip, _fp, _szp]),
    "wavecu_icp_aligned": (C.c_int, [_vp, _fp, _szp]),
    "wavecu_icp_trace": (C.c_int, [_vp, _dp, _ip, _fp, _ip]),
    "wavecu_icp_info": (C.c_int, [_vp, C.c_int, _dp]),
    "wavecu_icp_set_profiling": (C.c_int, [_vp, C.c_int]),
    "wavecu_icp_stats": (C.c_int, [_vp, C.POINTER(StatsC)]),
    "wavecu_nn_create": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "wavecu_nn_destroy": (C.c_int, [_vp]),
    "wavecu_nn_set_target": (C.c_int, [_vp, _fp, _sz]),
    "wavecu_nn_search": (C.c_int, [_vp, _fp, _sz, C.c_double, _ip, _fp]),
    "wavecu_nn_search_device": (C.c_int, [_vp, _vp, _sz, C.c_double, _vp, _vp, C.c_int, _fp]),
    "wavecu_icp_set_search": (C.c_int, [_vp, C.c_int]),
    "wavecu_icp_build_target": (C.c_int, [_vp]),
    "wavecu_icp_share_target": (C.c_int, [_vp, _vp]),
    "wavecu_batch_create": (C.c_int, [C.POINTER(IcpParamsC), _ip, C.c_int, C.c_int, C.POINTER(_vp)]),
    "wavecu_batch_destroy": (C.c_int, [_vp]),
    "wavecu_batch_device_count": (C.c_int, [_vp]),
    "wavecu_batch_set_map": (C.c_int, [_vp, _fp, _sz]),
    "wavecu_batch_match": (C.c_int, [_vp, C.POINTER(_fp), _szp, _ip, C.c_int, C.POINTER(_fp), _szp, C.c_int,
                                     C.POINTER(BatchRecordC)]),
    "wavecu_batch_unique_id": (C.c_int, [_vp]),
    "wavecu_batch_init_comm": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "wavecu_batch_broadcast_map": (C.c_int, [_vp, _fp, _sz, C.c_int]),
    "wavecu_batch_allgather": (C.c_int, [_vp, C.POINTER(BatchRecordC), C.c_int, C.POINTER(BatchRecordC)]),
    "wavecu_ndt_default_params": (None, [C.POINTER(NdtParamsC)]),
    "wavecu_ndt_create": (C.c_int, [C.POINTER(NdtParamsC), C.c_int, _vp, C.POINTER(_vp)]),
    "wavecu_ndt_destroy": (C.c_int, [_vp]),
    "wavecu_ndt_set_params": (C.c_int, [_vp, C.POINTER(NdtParamsC)]),
    "wavecu_ndt_set_source": (C.c_int, [_vp, _fp, _sz]),
    "wavecu_ndt_set_target": (C.c_int, [_vp, _fp, _sz]),
    "wavecu_ndt_set_source_device": (C.c_int, [_vp, _vp, _sz]),
    "wavecu_ndt_set_target_device": (C.c_int, [_vp, _vp, _sz]),
    "wavecu_ndt_match": (C.c_int, [_vp, _dp, _ip, _ip]),
    "wavecu_ndt_grid": (C.c_int, [_vp, _ip, _ip, _ip, _fp, _dp, _dp, C.c_int]),
    "wavecu_ndt_derivatives": (C.c_int, [_vp, _dp, _fp, _dp, _dp, _dp]),
    "wavecu_ndt_stats": (C.c_int, [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), _ip]),
    "wavecu_ndt_set_profiling": (C.c_int, [_vp, C.c_int]),
    "wavecu_ndt_timing": (C.c_int, [_vp, _dp]),
    "wavecu_gicp_set_profiling": (C.c_int, [_vp, C.c_int]),
    "wavecu_gicp_timing": (C.c_int, [_vp, _dp, C.POINTER(C.c_longlong)]),
    "wavecu_gicp_default_params": (None, [C.POINTER(GicpParamsC)]),
    "wavecu_gicp_create": (C.c_int, [C.POINTER(GicpParamsC), C.c_int, _vp, C.POINTER(_vp)]),
    "wavecu_gicp_destroy": (C.c_int, [_vp]),
    "wavecu_gicp_set_params": (C.c_int, [_vp, C.POINTER(GicpParamsC)]),
    "wavecu_gicp_set_source": (C.c_int, [_vp, _fp, _sz]),
    "wavecu_gicp_set_target": (C.c_int, [_vp, _fp, _sz]),
    "wavecu_gicp_set_source_device": (C.c_int, [_vp, _vp, _sz]),
    "wavecu_gicp_set_target_device": (C.c_int, [_vp, _vp, _sz]),
    "wavecu_gicp_match": (C.c_int, [_vp, _dp, _ip, _ip]),
    "wavecu_gicp_covariances": (C.c_int, [_vp, C.c_int, _dp, _szp]),
    "wavecu_gicp_cloud": (C.c_int, [_vp, C.c_int, _fp, _szp]),
    "wavecu_gicp_stats": (C.c_int, [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                    _szp]),
    "wavecu_voxel_grid": (C.c_int, [C.c_int, _fp, _sz, C.c_float, _fp, _szp, _ip]),
    "wavecu_last_error": (C.c_char_p, []),
    "wavecu_device_count": (C.c_int, []),
}

_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (python -m libwave_b200.build)."""
    global _lib
    if _lib is None:
        if not SO.exists():
            raise WavecuError(f"{SO} is missing - build it with __graft_entry__.build() / "
                              "python -m libwave_b200.build; there is no CPU fallback")
        L = C.CDLL(str(SO))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise WavecuError(f"wavecu error {rc}: {lib().wavecu_last_error().decode()}")
