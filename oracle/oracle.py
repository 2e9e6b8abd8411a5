"""ORACLE - test infrastructure only.

ctypes front end of oracle/_build/libwave_oracle.so (the CPU restatement of the PCL algorithms
the reference's match() calls; see oracle/*.hpp for the reference file:line each piece follows).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package libwave_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

_DIR = pathlib.Path(__file__).resolve().parent
_SO = _DIR / "_build" / "libwave_oracle.so"

EST_SVD, EST_POINT_TO_PLANE = 0, 1
SUM_EXACT, SUM_PCL = 0, 1
CONV_STATES = ("NOT_CONVERGED", "ITERATIONS", "TRANSFORM", "ABS_MSE", "REL_MSE", "NO_CORRESPONDENCES")


def build(force: bool = False) -> pathlib.Path:
    """Compile the oracle with g++ (make -C oracle).  Building the checker is not using it."""
    srcs = list(_DIR.glob("*.cpp")) + list(_DIR.glob("*.hpp")) + [_DIR / "Makefile"]
    stale = (not _SO.exists()) or any(s.stat().st_mtime > _SO.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", str(_DIR)], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not _SO.exists():
            build()
        _lib = C.CDLL(str(_SO))
        _declare(_lib)
    return _lib


class IcpParamsC(C.Structure):
    _fields_ = [("max_corr", C.c_double), ("max_iter", C.c_int), ("t_eps", C.c_double),
                ("fit_eps", C.c_double), ("estimator", C.c_int), ("sum_mode", C.c_int)]


_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


def _declare(L):
    L.wo_kdtree_create.restype = C.c_void_p
    L.wo_kdtree_create.argtypes = [_fp, C.c_size_t]
    L.wo_kdtree_destroy.argtypes = [C.c_void_p]
    L.wo_kdtree_nn1.argtypes = [C.c_void_p, _fp, C.c_size_t, _ip, _fp, C.c_int]
    L.wo_kdtree_knn.argtypes = [C.c_void_p, _fp, C.c_size_t, C.c_int, _ip, _fp]
    L.wo_brute_nn1.argtypes = [_fp, C.c_size_t, _fp, C.c_size_t, _ip, _fp]
    L.wo_icp_run.restype = C.c_void_p
    L.wo_icp_run.argtypes = [_fp, C.c_size_t, _fp, C.c_size_t, _fp, C.POINTER(IcpParamsC), C.c_void_p, C.c_int]
    L.wo_icp_result_summary.argtypes = [C.c_void_p, _fp, _ip, _ip, _ip, C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.wo_icp_result_corr.argtypes = [C.c_void_p, _ip, _ip, _fp]
    L.wo_icp_result_aligned.argtypes = [C.c_void_p, _fp]
    L.wo_icp_result_trace.argtypes = [C.c_void_p, _dp, _ip, _fp]
    L.wo_icp_result_free.argtypes = [C.c_void_p]
    L.wo_fix_scales.argtypes = [_fp, C.c_size_t, _fp, C.c_size_t, C.c_double, _ip]
    L.wo_rotation_from_sigma.argtypes = [_dp, _dp]
    L.wo_solve6.argtypes = [_dp, _dp, _dp]
    L.wo_solve6.restype = C.c_int


def xyzw(pts) -> np.ndarray:
    """(n,3) or (n,4) -> contiguous fp32 (n,4) with w = 1 (pcl::PointXYZ layout)."""
    pts = np.asarray(pts, dtype=np.float32)
    if pts.ndim == 2 and pts.shape[1] == 4:
        return np.ascontiguousarray(pts)
    out = np.ones((pts.shape[0], 4), dtype=np.float32)
    out[:, :3] = pts
    return out


def _f(a):
    return a.ctypes.data_as(_fp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _d(a):
    return a.ctypes.data_as(_dp)


class KdTree:
    def __init__(self, target):
        self.pts = xyzw(target)
        self.h = lib().wo_kdtree_create(_f(self.pts), self.pts.shape[0])

    def __del__(self):
        if getattr(self, "h", None):
            lib().wo_kdtree_destroy(self.h)
            self.h = None

    def nn1(self, queries, nthreads: int = 1):
        q = xyzw(queries)
        idx = np.empty(q.shape[0], dtype=np.int32)
        d2 = np.empty(q.shape[0], dtype=np.float32)
        lib().wo_kdtree_nn1(self.h, _f(q), q.shape[0], _i(idx), _f(d2), nthreads)
        return idx, d2

    def knn(self, queries, k: int):
        q = xyzw(queries)
        idx = np.empty((q.shape[0], k), dtype=np.int32)
        d2 = np.empty((q.shape[0], k), dtype=np.float32)
        lib().wo_kdtree_knn(self.h, _f(q), q.shape[0], k, _i(idx), _f(d2))
        return idx, d2


def brute_nn1(target, queries):
    t, q = xyzw(target), xyzw(queries)
    idx = np.empty(q.shape[0], dtype=np.int32)
    d2 = np.empty(q.shape[0], dtype=np.float32)
    lib().wo_brute_nn1(_f(t), t.shape[0], _f(q), q.shape[0], _i(idx), _f(d2))
    return idx, d2


class IcpResult:
    pass


def icp_align(source, target, *, max_corr=3.0, max_iter=100, t_eps=1e-8, fit_eps=1e-2,
              estimator=EST_SVD, sum_mode=SUM_EXACT, target_normals=None, tree: KdTree | None = None,
              nn_threads: int = 1) -> IcpResult:
    """pcl::IterativeClosestPoint::align restated (oracle/icp.cpp); defaults = icp.hpp:35-43."""
    s, t = xyzw(source), xyzw(target)
    nrm = xyzw(target_normals) if target_normals is not None else None
    prm = IcpParamsC(max_corr, max_iter, t_eps, fit_eps, estimator, sum_mode)
    L = lib()
    h = L.wo_icp_run(_f(s), s.shape[0], _f(t), t.shape[0], _f(nrm) if nrm is not None else None,
                     C.byref(prm), tree.h if tree is not None else None, nn_threads)
    T = np.empty(16, dtype=np.float32)
    conv, iters, state = C.c_int(), C.c_int(), C.c_int()
    nc, nt, na = C.c_size_t(), C.c_size_t(), C.c_size_t()
    L.wo_icp_result_summary(h, _f(T), C.byref(conv), C.byref(iters), C.byref(state), C.byref(nc),
                            C.byref(nt), C.byref(na))
    r = IcpResult()
    r.T = T.reshape(4, 4)
    r.converged, r.iterations, r.state = bool(conv.value), iters.value, CONV_STATES[state.value]
    r.corr_query = np.empty(nc.value, dtype=np.int32)
    r.corr_match = np.empty(nc.value, dtype=np.int32)
    r.corr_dist = np.empty(nc.value, dtype=np.float32)
    L.wo_icp_result_corr(h, _i(r.corr_query), _i(r.corr_match), _f(r.corr_dist))
    r.aligned = np.empty((na.value, 4), dtype=np.float32)
    L.wo_icp_result_aligned(h, _f(r.aligned))
    r.mse = np.empty(nt.value, dtype=np.float64)
    r.n_corr = np.empty(nt.value, dtype=np.int32)
    r.T_trace = np.empty((nt.value, 4, 4), dtype=np.float32)
    L.wo_icp_result_trace(h, _d(r.mse), _i(r.n_corr), _f(r.T_trace))
    L.wo_icp_result_free(h)
    return r


def fix_scales(source, target, max_corr):
    s, t = xyzw(source), xyzw(target)
    k = np.empty(3, dtype=np.int32)
    lib().wo_fix_scales(_f(s), s.shape[0], _f(t), t.shape[0], max_corr, _i(k))
    return tuple(int(v) for v in k)


def fix128_sum(terms, k: int) -> float:
    """Exact 128-bit fixed-point sum of round(terms * 2^k), converted back to fp64 (smallmat.hpp)."""
    t = np.ascontiguousarray(terms, dtype=np.float64)
    L = lib()
    L.wo_fix128_sum.restype = C.c_double
    L.wo_fix128_sum.argtypes = [_dp, C.c_size_t, C.c_int]
    return float(L.wo_fix128_sum(_d(t), t.shape[0], k))


def rotation_from_sigma(S):
    S = np.ascontiguousarray(S, dtype=np.float64).reshape(9)
    R = np.empty(9, dtype=np.float64)
    lib().wo_rotation_from_sigma(_d(S), _d(R))
    return R.reshape(3, 3)


def solve6(A, b):
    A = np.ascontiguousarray(A, dtype=np.float64).reshape(36)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(6)
    x = np.empty(6, dtype=np.float64)
    ok = lib().wo_solve6(_d(A), _d(b), _d(x))
    return x if ok else None


# ---- matcher level (oracle/matcher.cpp) -----------------------------------------------------------
class MatcherParamsC(C.Structure):
    _fields_ = [("max_corr", C.c_double), ("max_iter", C.c_int), ("t_eps", C.c_double), ("fit_eps", C.c_double),
                ("multiscale_steps", C.c_int), ("res", C.c_float), ("sum_mode", C.c_int)]


def _declare_matcher(L):
    L.wo_voxel_grid.restype = C.c_size_t
    L.wo_voxel_grid.argtypes = [_fp, C.c_size_t, C.c_float, _fp, _ip]
    L.wo_transform_affine3d.argtypes = [_fp, C.c_size_t, _dp, _fp]
    L.wo_icp_match.restype = C.c_void_p
    L.wo_icp_match.argtypes = [_fp, C.c_size_t, _fp, C.c_size_t, C.POINTER(MatcherParamsC), C.c_int, _ip]
    L.wo_match_summary.argtypes = [C.c_void_p, _dp, _ip, _ip, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.wo_match_clouds.argtypes = [C.c_void_p, _fp, _fp]
    L.wo_match_last.restype = C.c_void_p
    L.wo_match_last.argtypes = [C.c_void_p]
    L.wo_match_free.argtypes = [C.c_void_p]
    L.wo_estimate_lum.restype = C.c_int
    L.wo_estimate_lum.argtypes = [_fp, _fp, _ip, _ip, C.c_size_t, C.c_int, C.c_int, C.c_int, _dp]
    L.wo_estimate_lum_old.restype = C.c_int
    L.wo_estimate_lum_old.argtypes = [_fp, C.c_size_t, _fp, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_int, _dp,
                                      C.c_int]


_orig_declare = _declare


def _declare(L):  # noqa: F811
    _orig_declare(L)
    _declare_matcher(L)


def voxel_grid(cloud, leaf: float):
    """pcl::VoxelGrid<PointXYZ>::filter; returns (xyzw, filtered_flag)."""
    a = xyzw(cloud)
    out = np.empty_like(a)
    flag = C.c_int()
    n = lib().wo_voxel_grid(_f(a), a.shape[0], leaf, _f(out), C.byref(flag))
    return out[:n].copy(), bool(flag.value)


def transform_affine3d(cloud, T):
    a = xyzw(cloud)
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
    out = np.empty_like(a)
    lib().wo_transform_affine3d(_f(a), a.shape[0], _d(T), _f(out))
    return out


def _read_icp_result(h) -> IcpResult:
    L = lib()
    T = np.empty(16, dtype=np.float32)
    conv, iters, state = C.c_int(), C.c_int(), C.c_int()
    nc, nt, na = C.c_size_t(), C.c_size_t(), C.c_size_t()
    L.wo_icp_result_summary(h, _f(T), C.byref(conv), C.byref(iters), C.byref(state), C.byref(nc), C.byref(nt),
                            C.byref(na))
    r = IcpResult()
    r.T = T.reshape(4, 4)
    r.converged, r.iterations, r.state = bool(conv.value), iters.value, CONV_STATES[state.value]
    r.corr_query = np.empty(nc.value, dtype=np.int32)
    r.corr_match = np.empty(nc.value, dtype=np.int32)
    r.corr_dist = np.empty(nc.value, dtype=np.float32)
    L.wo_icp_result_corr(h, _i(r.corr_query), _i(r.corr_match), _f(r.corr_dist))
    r.aligned = np.empty((na.value, 4), dtype=np.float32)
    L.wo_icp_result_aligned(h, _f(r.aligned))
    r.mse = np.empty(nt.value, dtype=np.float64)
    r.n_corr = np.empty(nt.value, dtype=np.int32)
    r.T_trace = np.empty((nt.value, 4, 4), dtype=np.float32)
    L.wo_icp_result_trace(h, _d(r.mse), _i(r.n_corr), _f(r.T_trace))
    return r


class MatchResult:
    pass


def icp_match(ref, target, *, max_corr=3.0, max_iter=100, t_eps=1e-8, fit_eps=1e-2, multiscale_steps=3, res=0.1,
              sum_mode=SUM_EXACT, nn_threads: int = 1) -> MatchResult:
    """wave::ICPMatcher::match() restated (src/icp.cpp:75-133); defaults = icp.hpp:35-59."""
    s, t = xyzw(ref), xyzw(target)
    prm = MatcherParamsC(max_corr, max_iter, t_eps, fit_eps, multiscale_steps, res, sum_mode)
    ok = C.c_int()
    L = lib()
    h = L.wo_icp_match(_f(s), s.shape[0], _f(t), t.shape[0], C.byref(prm), nn_threads, C.byref(ok))
    T = np.empty(16, dtype=np.float64)
    lv, it = C.c_int(), C.c_int()
    nr, ntg = C.c_size_t(), C.c_size_t()
    L.wo_match_summary(h, _d(T), C.byref(lv), C.byref(it), C.byref(nr), C.byref(ntg))
    r = MatchResult()
    r.success, r.T, r.levels, r.total_iterations = bool(ok.value), T.reshape(4, 4), lv.value, it.value
    r.ds_ref = np.empty((nr.value, 4), dtype=np.float32)
    r.ds_tgt = np.empty((ntg.value, 4), dtype=np.float32)
    L.wo_match_clouds(h, _f(r.ds_ref), _f(r.ds_tgt))
    r.last = _read_icp_result(L.wo_match_last(h))
    L.wo_match_free(h)
    return r


def estimate_lum(aligned, target, corr_q, corr_m, sum_mode=SUM_EXACT, k_quad: int = 42):
    """estimateLUM; k_quad = fix_scales(source, target, max_corr)[1] of the align() it follows."""
    a, t = xyzw(aligned), xyzw(target)
    q = np.ascontiguousarray(corr_q, dtype=np.int32)
    m = np.ascontiguousarray(corr_m, dtype=np.int32)
    info = np.empty(36, dtype=np.float64)
    ok = lib().wo_estimate_lum(_f(a), _f(t), _i(q), _i(m), q.shape[0], sum_mode, k_quad - 2, k_quad - 4, _d(info))
    return info.reshape(6, 6), bool(ok)


def estimate_lum_old(aligned, target, max_corr, sum_mode=SUM_EXACT, k_quad: int = 42, nn_threads: int = 1):
    a, t = xyzw(aligned), xyzw(target)
    info = np.empty(36, dtype=np.float64)
    ok = lib().wo_estimate_lum_old(_f(a), a.shape[0], _f(t), t.shape[0], max_corr, sum_mode, k_quad - 2, k_quad - 4,
                                   _d(info), nn_threads)
    return info.reshape(6, 6), bool(ok)


def estimate_censi(ref, target, corr_q, corr_m, T, lin_covar=2.5e-4, ang_covar=7.78e-9):
    """ICPMatcher::estimateCensi (oracle/censi.cpp).  Returns (information, d2J_dX2, middle, ok)."""
    r, t = xyzw(ref), xyzw(target)
    q = np.ascontiguousarray(corr_q, dtype=np.int32)
    m = np.ascontiguousarray(corr_m, dtype=np.int32)
    T16 = np.ascontiguousarray(np.asarray(T, dtype=np.float64).reshape(16))
    H, M, info = (np.empty(36, dtype=np.float64) for _ in range(3))
    L = lib()
    L.wo_estimate_censi.argtypes = [_fp, _fp, _ip, _ip, C.c_size_t, _dp, C.c_double, C.c_double, _dp, _dp, _dp]
    L.wo_estimate_censi.restype = C.c_int
    ok = L.wo_estimate_censi(_f(r), _f(t), _i(q), _i(m), q.shape[0], _d(T16), lin_covar, ang_covar, _d(H), _d(M),
                             _d(info))
    return info.reshape(6, 6), H.reshape(6, 6), M.reshape(6, 6), bool(ok)


# ---- NDT (oracle/ndt.cpp) -------------------------------------------------------------------------
class NdtParamsC(C.Structure):
    _fields_ = [("step_size", C.c_int), ("max_iter", C.c_int), ("t_eps", C.c_double), ("res", C.c_float),
                ("line_search", C.c_int)]


class NdtResult:
    pass


NDT_LS_PCL18, NDT_LS_MORE_THUENTE = 0, 1


def ndt_align(source, target, *, step_size=3, max_iter=100, t_eps=1e-8, res=5.0,
              line_search=NDT_LS_MORE_THUENTE) -> NdtResult:
    """pcl::NormalDistributionsTransform::align as NDTMatcher drives it (src/ndt.cpp:18-65).
    line_search: NDT_LS_MORE_THUENTE (PCL >= 1.9) or NDT_LS_PCL18 (the search PCL 1.8 skips)."""
    s, t = xyzw(source), xyzw(target)
    L = lib()
    L.wo_ndt_align.argtypes = [_fp, C.c_size_t, _fp, C.c_size_t, C.POINTER(NdtParamsC), _fp, _dp, _ip, _ip, _ip, _dp,
                               _dp, _ip]
    prm = NdtParamsC(step_size, max_iter, t_eps, res, line_search)
    T = np.empty(16, dtype=np.float32)
    pose = np.empty(6, dtype=np.float64)
    conv, iters, nv, ntr = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    score = C.c_double()
    trace = np.empty(max_iter + 4, dtype=np.float64)
    L.wo_ndt_align(_f(s), s.shape[0], _f(t), t.shape[0], C.byref(prm), _f(T), _d(pose), C.byref(conv), C.byref(iters),
                   C.byref(nv), C.byref(score), _d(trace), C.byref(ntr))
    r = NdtResult()
    r.T, r.pose, r.converged, r.iterations = T.reshape(4, 4), pose, bool(conv.value), iters.value
    r.n_voxels, r.score, r.steps = nv.value, score.value, trace[:ntr.value].copy()
    return r


def ndt_grid(target, res):
    t = xyzw(target)
    n = t.shape[0]
    L = lib()
    L.wo_ndt_grid.restype = C.c_size_t
    L.wo_ndt_grid.argtypes = [_fp, C.c_size_t, C.c_float, _ip, _ip, _fp, _dp, _dp]
    voxel = np.empty(n, np.int32)
    count = np.empty(n, np.int32)
    cen = np.empty((n, 3), np.float32)
    mean = np.empty((n, 3), np.float64)
    icov = np.empty((n, 9), np.float64)
    m = L.wo_ndt_grid(_f(t), n, res, _i(voxel), _i(count), _f(cen), _d(mean), _d(icov))
    return voxel[:m].copy(), count[:m].copy(), cen[:m].copy(), mean[:m].copy(), icov[:m].reshape(m, 3, 3).copy()


def ndt_derivatives(source, target, res, pose, T):
    s, t = xyzw(source), xyzw(target)
    L = lib()
    L.wo_ndt_derivatives.restype = C.c_double
    L.wo_ndt_derivatives.argtypes = [_fp, C.c_size_t, _fp, C.c_size_t, C.c_float, _dp, _fp, _dp, _dp]
    pose = np.ascontiguousarray(pose, dtype=np.float64)
    T = np.ascontiguousarray(T, dtype=np.float32).reshape(16)
    g = np.empty(6, np.float64)
    H = np.empty(36, np.float64)
    score = L.wo_ndt_derivatives(_f(s), s.shape[0], _f(t), t.shape[0], res, _d(pose), _f(T), _d(g), _d(H))
    return score, g, H.reshape(6, 6)


# ---- GICP (oracle/gicp.cpp) -----------------------------------------------------------------------
class GicpResult:
    pass


def gicp_covariances(cloud, k=10, eps=1e-3):
    c = xyzw(cloud)
    L = lib()
    L.wo_gicp_covariances.argtypes = [_fp, C.c_size_t, C.c_int, C.c_double, _dp]
    out = np.empty((c.shape[0], 9), dtype=np.float64)
    L.wo_gicp_covariances(_f(c), c.shape[0], k, eps, _d(out))
    return out.reshape(-1, 3, 3)


def gicp_align(source, target, *, corr_rand=10, max_iter=100, r_eps=1e-8) -> GicpResult:
    """pcl::GeneralizedIterativeClosestPoint::align as GICPMatcher drives it (src/gicp.cpp:20-64);
    voxel filtering (src/gicp.cpp:37-55) is the caller's job: pass voxel_grid() outputs."""
    s, t = xyzw(source), xyzw(target)
    L = lib()
    L.wo_gicp_align.argtypes = [_fp, C.c_size_t, _fp, C.c_size_t, C.c_int, C.c_int, C.c_double, _fp, _ip, _ip,
                                C.POINTER(C.c_size_t), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), _dp, _ip]
    T = np.empty(16, dtype=np.float32)
    conv, iters, ntr = C.c_int(), C.c_int(), C.c_int()
    nc = C.c_size_t()
    inner, evals = C.c_longlong(), C.c_longlong()
    trace = np.empty(max_iter + 4, dtype=np.float64)
    L.wo_gicp_align(_f(s), s.shape[0], _f(t), t.shape[0], corr_rand, max_iter, r_eps, _f(T), C.byref(conv),
                    C.byref(iters), C.byref(nc), C.byref(inner), C.byref(evals), _d(trace), C.byref(ntr))
    r = GicpResult()
    r.T, r.converged, r.iterations, r.n_corr = T.reshape(4, 4), bool(conv.value), iters.value, nc.value
    r.inner_iterations, r.evaluations, r.delta = inner.value, evals.value, trace[:ntr.value].copy()
    return r


def estimate_normals(cloud, k=10):
    """k-NN PCA normals oriented towards the origin (the repo's point-to-plane extension)."""
    c = xyzw(cloud)
    L = lib()
    L.wo_estimate_normals.argtypes = [_fp, C.c_size_t, C.c_int, _fp]
    out = np.empty_like(c)
    L.wo_estimate_normals(_f(c), c.shape[0], k, _f(out))
    return out
