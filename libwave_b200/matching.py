"""Python mirror of the reference's matcher interface over the C ABI (include/wavecu.h).

Same names, argument meaning and error behaviour as wave_matching's C++ classes
(include/wave/matching/matcher.hpp:23-99, icp.hpp:30-120): ``setRef`` / ``setTarget`` / ``setup``
/ ``match`` / ``estimateInfo`` / ``getResult`` / ``getInfo`` / ``getRes``, a public ``params``
member, ``match()`` returning False on non-convergence and never raising for it.  Clouds are
(n,3) or (n,4) float32 arrays (pcl::PointXYZ records).  Everything numeric runs in libwavecu.so;
there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from . import capi
from .capi import EST_POINT_TO_PLANE, EST_SVD, INFO_CENSI, INFO_LUM, INFO_LUMOLD  # noqa: F401


def _xyzw(pts) -> np.ndarray:
    pts = np.asarray(pts, dtype=np.float32)
    if pts.ndim != 2 or pts.shape[1] not in (3, 4):
        raise ValueError("cloud must be (n,3) or (n,4) float32")
    if pts.shape[1] == 4:
        return np.ascontiguousarray(pts)
    out = np.ones((pts.shape[0], 4), dtype=np.float32)
    out[:, :3] = pts
    return out


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@dataclasses.dataclass
class ICPMatcherParams:
    """wave::ICPMatcherParams, defaults from icp.hpp:35-64."""
    max_corr: float = 3.0
    max_iter: int = 100
    t_eps: float = 1e-8
    fit_eps: float = 1e-2
    lidar_ang_covar: float = 7.78e-9
    lidar_lin_covar: float = 2.5e-4
    multiscale_steps: int = 3
    res: float = 0.1
    covar_estimator: int = INFO_LUM
    estimator: int = EST_SVD  # extension: EST_POINT_TO_PLANE

    def to_c(self) -> capi.IcpParamsC:
        return capi.IcpParamsC(self.max_corr, self.max_iter, self.t_eps, self.fit_eps, self.lidar_ang_covar,
                               self.lidar_lin_covar, self.multiscale_steps, self.res, self.covar_estimator,
                               self.estimator)


class Matcher:
    """wave::Matcher<T> (matcher.hpp:23-99)."""

    def __init__(self, res: float = -1.0):
        self.resolution = float(res)
        self.result = np.eye(4)
        self.information = np.eye(6)

    def getResult(self):
        return self.result.copy()

    def getInfo(self):
        return self.information

    def getRes(self):
        return self.resolution

    def setRef(self, ref):
        raise NotImplementedError

    def setTarget(self, target):
        raise NotImplementedError

    def setup(self, ref, target):
        self.setRef(ref)
        self.setTarget(target)

    def match(self) -> bool:
        return False

    def estimateInfo(self):
        self.information = np.eye(6)


class ICPMatcher(Matcher):
    """wave::ICPMatcher (icp.hpp:67-120, src/icp.cpp:32-142) on the GPU."""

    def __init__(self, params: ICPMatcherParams | None = None, device: int = 0, stream: int | None = None):
        self.params = dataclasses.replace(params) if params is not None else ICPMatcherParams()
        super().__init__(self.params.res)
        self._L = capi.lib()
        self._h = C.c_void_p()
        prm = self.params.to_c()
        capi.check(self._L.wavecu_icp_create(C.byref(prm), device, C.c_void_p(stream or 0), C.byref(self._h)))
        self._n_ref = 0
        self.converged = False
        self.iterations = 0

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.wavecu_icp_destroy(h)
            self._h = None

    # -- clouds ----------------------------------------------------------------------------------
    def setRef(self, ref):
        """(n, 3) or (n, 4) float32.  A page-locked (n, 4) array is handed to the library as it is and
        copied asynchronously: keep it alive and unmodified until match() has returned (wavecu.h)."""
        a = _xyzw(ref)
        self._n_ref = a.shape[0]
        self._keep_ref = a          # the upload may still be pending when this method returns
        capi.check(self._L.wavecu_icp_set_source(self._h, _f(a), a.shape[0]))

    def setTarget(self, target):
        a = _xyzw(target)
        self._keep_tgt = a
        capi.check(self._L.wavecu_icp_set_target(self._h, _f(a), a.shape[0]))

    def setTargetNormals(self, normals):
        a = _xyzw(normals)
        self._keep_nrm = a
        capi.check(self._L.wavecu_icp_set_target_normals(self._h, _f(a), a.shape[0]))

    def setRefDevice(self, ptr: int, n: int):
        self._n_ref = n
        capi.check(self._L.wavecu_icp_set_source_device(self._h, C.c_void_p(ptr), n))

    def setTargetDevice(self, ptr: int, n: int):
        capi.check(self._L.wavecu_icp_set_target_device(self._h, C.c_void_p(ptr), n))

    def setTargetNormalsDevice(self, ptr: int, n: int):
        capi.check(self._L.wavecu_icp_set_target_normals_device(self._h, C.c_void_p(ptr), n))

    # -- match -----------------------------------------------------------------------------------
    def _push_params(self):
        prm = self.params.to_c()
        capi.check(self._L.wavecu_icp_set_params(self._h, C.byref(prm)))

    def match(self) -> bool:
        self._push_params()
        T = np.empty(16, dtype=np.float64)
        conv, iters = C.c_int(), C.c_int()
        capi.check(self._L.wavecu_icp_match(self._h, _d(T), C.byref(conv), C.byref(iters)))
        self.converged, self.iterations = bool(conv.value), iters.value
        if self.converged:
            self.result = T.reshape(4, 4).copy()
            return True
        return False

    def align(self):
        """One pcl align() at the current resolution; returns (T, converged, iterations, state)."""
        self._push_params()
        T = np.empty(16, dtype=np.float64)
        conv, iters, state = C.c_int(), C.c_int(), C.c_int()
        capi.check(self._L.wavecu_icp_align(self._h, _d(T), C.byref(conv), C.byref(iters), C.byref(state)))
        self.converged, self.iterations = bool(conv.value), iters.value
        return T.reshape(4, 4).copy(), bool(conv.value), iters.value, capi.CONV_STATES[state.value]

    def info(self, method: int):
        """One estimator on its own: INFO_LUM (estimateLUM), INFO_CENSI (estimateCensi) or INFO_LUMOLD
        (estimateLUMold)."""
        info = np.empty(36, dtype=np.float64)
        capi.check(self._L.wavecu_icp_info(self._h, method, _d(info)))
        return info.reshape(6, 6).copy()

    def estimateInfo(self):
        """ICPMatcher::estimateInfo (src/icp.cpp:135-142).  The reference's switch has no breaks:
        LUM falls through to Censi and then LUMold, Censi falls through to LUMold - so whatever
        covar_estimator says, `information` ends up as estimateLUMold's result.  That observable
        behaviour is what is reproduced (the intermediate Censi matrix is never visible)."""
        if self.params.covar_estimator == INFO_LUM:
            self.information = self.info(INFO_LUM)
        if self.params.covar_estimator in (INFO_LUM, INFO_CENSI, INFO_LUMOLD):
            self.information = self.info(INFO_LUMOLD)

    # -- introspection ---------------------------------------------------------------------------
    def correspondences(self):
        n = C.c_size_t()
        q = np.empty(max(self._aligned_size(), 1), dtype=np.int32)
        m = np.empty_like(q)
        d2 = np.empty(q.shape[0], dtype=np.float32)
        capi.check(self._L.wavecu_icp_correspondences(self._h, _i(q), _i(m), _f(d2), C.byref(n)))
        return q[:n.value].copy(), m[:n.value].copy(), d2[:n.value].copy()

    def _aligned_size(self) -> int:
        n = C.c_size_t()
        capi.check(self._L.wavecu_icp_aligned(self._h, None, C.byref(n)))
        return n.value

    def aligned(self):
        n = self._aligned_size()
        out = np.empty((n, 4), dtype=np.float32)
        nn = C.c_size_t()
        capi.check(self._L.wavecu_icp_aligned(self._h, _f(out), C.byref(nn)))
        return out

    def trace(self):
        cap = max(self.params.max_iter, 1)
        mse = np.empty(cap, dtype=np.float64)
        nc = np.empty(cap, dtype=np.int32)
        T = np.empty((cap, 4, 4), dtype=np.float32)
        n = C.c_int()
        capi.check(self._L.wavecu_icp_trace(self._h, _d(mse), _i(nc), _f(T), C.byref(n)))
        return mse[:n.value].copy(), nc[:n.value].copy(), T[:n.value].copy()

    def set_profiling(self, on: bool):
        capi.check(self._L.wavecu_icp_set_profiling(self._h, int(on)))

    def set_search(self, mode: int):
        """SEARCH_TREE (LBVH walk, default) or SEARCH_TILED (shared-memory tiles, csrc/tile_nn.cuh)."""
        capi.check(self._L.wavecu_icp_set_search(self._h, int(mode)))

    def stats(self) -> dict:
        s = capi.StatsC()
        capi.check(self._L.wavecu_icp_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in capi.StatsC._fields_}


@dataclasses.dataclass
class GICPMatcherParams:
    """wave::GICPMatcherParams, defaults from gicp.hpp:34-38."""
    corr_rand: int = 10
    max_iter: int = 100
    r_eps: float = 1e-8
    fit_eps: float = 1e-2
    res: float = 0.1

    def to_c(self) -> capi.GicpParamsC:
        return capi.GicpParamsC(self.corr_rand, self.max_iter, self.r_eps, self.fit_eps, self.res)


class GICPMatcher(Matcher):
    """wave::GICPMatcher (gicp.hpp:41-65, src/gicp.cpp:20-64) on the GPU.  As in the reference the
    voxel filter runs inside setRef / setTarget and the parameters are fixed at construction."""

    def __init__(self, params: GICPMatcherParams | None = None, device: int = 0, stream: int | None = None):
        self.params = dataclasses.replace(params) if params is not None else GICPMatcherParams()
        super().__init__(self.params.res if self.params.res > 0 else -1.0)
        self._L = capi.lib()
        self._h = C.c_void_p()
        prm = self.params.to_c()
        capi.check(self._L.wavecu_gicp_create(C.byref(prm), device, C.c_void_p(stream or 0), C.byref(self._h)))
        self.converged = False
        self.iterations = 0

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.wavecu_gicp_destroy(h)
            self._h = None

    def setRef(self, ref):
        a = _xyzw(ref)
        capi.check(self._L.wavecu_gicp_set_source(self._h, _f(a), a.shape[0]))

    def setTarget(self, target):
        a = _xyzw(target)
        capi.check(self._L.wavecu_gicp_set_target(self._h, _f(a), a.shape[0]))

    def setRefDevice(self, ptr: int, n: int):
        capi.check(self._L.wavecu_gicp_set_source_device(self._h, C.c_void_p(ptr), n))

    def setTargetDevice(self, ptr: int, n: int):
        capi.check(self._L.wavecu_gicp_set_target_device(self._h, C.c_void_p(ptr), n))

    def match(self) -> bool:
        T = np.empty(16, dtype=np.float64)
        conv, iters = C.c_int(), C.c_int()
        capi.check(self._L.wavecu_gicp_match(self._h, _d(T), C.byref(conv), C.byref(iters)))
        self.converged, self.iterations = bool(conv.value), iters.value
        if self.converged:
            self.result = T.reshape(4, 4).copy()
            return True
        return False

    def covariances(self, which: int):
        """(filtered cloud, per-point 3x3 covariances); which = 0 source, 1 target."""
        n = C.c_size_t()
        capi.check(self._L.wavecu_gicp_covariances(self._h, which, None, C.byref(n)))
        covs = np.empty((n.value, 9), dtype=np.float64)
        cloud = np.empty((n.value, 4), dtype=np.float32)
        capi.check(self._L.wavecu_gicp_covariances(self._h, which, _d(covs), C.byref(n)))
        capi.check(self._L.wavecu_gicp_cloud(self._h, which, _f(cloud), C.byref(n)))
        return cloud, covs.reshape(-1, 3, 3)

    def set_profiling(self, on: bool):
        capi.check(self._L.wavecu_gicp_set_profiling(self._h, int(on)))

    def stats(self) -> dict:
        a, b, c = C.c_longlong(), C.c_longlong(), C.c_longlong()
        n = C.c_size_t()
        capi.check(self._L.wavecu_gicp_stats(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(n)))
        ms, k = C.c_double(), C.c_longlong()
        capi.check(self._L.wavecu_gicp_timing(self._h, C.byref(ms), C.byref(k)))
        return {"kernel_launches": a.value, "evaluations": b.value, "inner_iterations": c.value, "n_corr": n.value,
                "cost_kernel_ms": ms.value, "cost_kernel_launches": k.value}


@dataclasses.dataclass
class NDTMatcherParams:
    """wave::NDTMatcherParams, defaults from ndt.hpp:37-41 (step_size is an int in the reference)."""
    step_size: int = 3
    max_iter: int = 100
    t_eps: float = 1e-8
    res: float = 5.0
    min_res: float = 0.05
    # not a reference field (wavecu.h): NDT_LS_MORE_THUENTE = PCL >= 1.9's computeStepLengthMT,
    # NDT_LS_PCL18 = PCL 1.8's, whose More-Thuente search never runs
    line_search: int = 1

    def to_c(self) -> capi.NdtParamsC:
        return capi.NdtParamsC(int(self.step_size), self.max_iter, self.t_eps, self.res, int(self.line_search))


NDT_LS_PCL18, NDT_LS_MORE_THUENTE = 0, 1
SEARCH_TREE, SEARCH_TILED = 0, 1


class NDTMatcher(Matcher):
    """wave::NDTMatcher (ndt.hpp:44-80, src/ndt.cpp:18-65) on the GPU."""

    def __init__(self, params: NDTMatcherParams | None = None, device: int = 0, stream: int | None = None):
        self.params = dataclasses.replace(params) if params is not None else NDTMatcherParams()
        if self.params.res < self.params.min_res:  # src/ndt.cpp:23-26
            print("[ERROR] Invalid resolution given, using minimum")
            self.params.res = self.params.min_res
        super().__init__(self.params.res)
        self._L = capi.lib()
        self._h = C.c_void_p()
        prm = self.params.to_c()
        capi.check(self._L.wavecu_ndt_create(C.byref(prm), device, C.c_void_p(stream or 0), C.byref(self._h)))
        self.converged = False
        self.iterations = 0

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.wavecu_ndt_destroy(h)
            self._h = None

    def setRef(self, ref):
        a = _xyzw(ref)
        capi.check(self._L.wavecu_ndt_set_source(self._h, _f(a), a.shape[0]))

    def setTarget(self, target):
        a = _xyzw(target)
        capi.check(self._L.wavecu_ndt_set_target(self._h, _f(a), a.shape[0]))

    def setRefDevice(self, ptr: int, n: int):
        capi.check(self._L.wavecu_ndt_set_source_device(self._h, C.c_void_p(ptr), n))

    def setTargetDevice(self, ptr: int, n: int):
        capi.check(self._L.wavecu_ndt_set_target_device(self._h, C.c_void_p(ptr), n))

    def match(self) -> bool:
        prm = self.params.to_c()
        capi.check(self._L.wavecu_ndt_set_params(self._h, C.byref(prm)))
        T = np.empty(16, dtype=np.float64)
        conv, iters = C.c_int(), C.c_int()
        capi.check(self._L.wavecu_ndt_match(self._h, _d(T), C.byref(conv), C.byref(iters)))
        self.converged, self.iterations = bool(conv.value), iters.value
        if self.converged:
            self.result = T.reshape(4, 4).copy()
            return True
        return False

    def grid(self):
        """(voxel index, count, fp32 centroid, fp64 mean, fp64 inverse covariance) of every cell."""
        prm = self.params.to_c()
        capi.check(self._L.wavecu_ndt_set_params(self._h, C.byref(prm)))
        n = C.c_int()
        capi.check(self._L.wavecu_ndt_grid(self._h, C.byref(n), None, None, None, None, None, 0))
        m = n.value
        voxel, count = np.empty(m, np.int32), np.empty(m, np.int32)
        cen, mean, icov = np.empty((m, 3), np.float32), np.empty((m, 3), np.float64), np.empty((m, 9), np.float64)
        capi.check(self._L.wavecu_ndt_grid(self._h, C.byref(n), _i(voxel), _i(count), _f(cen), _d(mean), _d(icov), m))
        return voxel, count, cen, mean, icov.reshape(m, 3, 3)

    def derivatives(self, pose, T):
        prm = self.params.to_c()
        capi.check(self._L.wavecu_ndt_set_params(self._h, C.byref(prm)))
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        T = np.ascontiguousarray(T, dtype=np.float32).reshape(16)
        score = C.c_double()
        g, H = np.empty(6, np.float64), np.empty(36, np.float64)
        capi.check(self._L.wavecu_ndt_derivatives(self._h, _d(pose), _f(T), C.byref(score), _d(g), _d(H)))
        return score.value, g, H.reshape(6, 6)

    def set_profiling(self, on: bool):
        capi.check(self._L.wavecu_ndt_set_profiling(self._h, int(on)))

    def stats(self) -> dict:
        a, b, c = C.c_longlong(), C.c_longlong(), C.c_int()
        capi.check(self._L.wavecu_ndt_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        ms = C.c_double()
        capi.check(self._L.wavecu_ndt_timing(self._h, C.byref(ms)))
        return {"kernel_launches": a.value, "derivative_passes": b.value, "n_cells": c.value,
                "derivative_kernel_ms": ms.value}


def voxel_grid(cloud, leaf: float, device: int = 0):
    """pcl::VoxelGrid<pcl::PointXYZ>::filter on the GPU; returns (xyzw, filtered)."""
    a = _xyzw(cloud)
    out = np.empty_like(a)
    n, flag = C.c_size_t(), C.c_int()
    capi.check(capi.lib().wavecu_voxel_grid(device, _f(a), a.shape[0], leaf, _f(out), C.byref(n), C.byref(flag)))
    return out[:n.value].copy(), bool(flag.value)


class NearestNeighbour:
    """Exact 1-NN over a target cloud (pcl::KdTreeFLANN nearestKSearch(k=1) semantics; lowest index
    among exact fp32 distance ties)."""

    def __init__(self, target=None, device: int = 0, stream: int | None = None):
        self._L = capi.lib()
        self._h = C.c_void_p()
        capi.check(self._L.wavecu_nn_create(device, C.c_void_p(stream or 0), C.byref(self._h)))
        if target is not None:
            self.set_target(target)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.wavecu_nn_destroy(h)
            self._h = None

    def set_target(self, target):
        a = _xyzw(target)
        capi.check(self._L.wavecu_nn_set_target(self._h, _f(a), a.shape[0]))

    def search(self, queries, max_dist: float = 0.0):
        q = _xyzw(queries)
        idx = np.empty(q.shape[0], dtype=np.int32)
        d2 = np.empty(q.shape[0], dtype=np.float32)
        capi.check(self._L.wavecu_nn_search(self._h, _f(q), q.shape[0], max_dist, _i(idx), _f(d2)))
        return idx, d2

    def search_device(self, q_ptr: int, nq: int, idx_ptr: int, d2_ptr: int, max_dist: float = 0.0, repeats: int = 1):
        ms = C.c_float()
        capi.check(self._L.wavecu_nn_search_device(self._h, C.c_void_p(q_ptr), nq, max_dist, C.c_void_p(idx_ptr),
                                                   C.c_void_p(d2_ptr), repeats, C.byref(ms)))
        return ms.value
