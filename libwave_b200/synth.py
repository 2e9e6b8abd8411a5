"""Synthetic Velodyne-style scan generator shared by the tests, the bench and the CPU baseline.

Spec: SURVEY.md section 8(d).  A spinning lidar at the sensor origin casts R rings x A azimuth
steps against a fixed scene (ground plane z = -1.73 m, the four walls of a 120 m x 80 m box and
64 vertical cylinders), adds Gaussian range noise and returns fp32 points in the *sensor* frame.

Convention (reference tests, wave_matching/tests/icp_tests.cpp:31-32,59): ``target = T_true * ref``
and ``match()`` is expected to return ``T_true``; a target scan is therefore taken from sensor
pose ``T_true^-1``.
"""
from __future__ import annotations

import numpy as np

GROUND_Z = -1.73
BOX_HALF = (60.0, 40.0)
N_CYL = 64
MAX_RANGE = 120.0
SIGMA = 0.02
ELEV_DEG = (-24.8, 2.0)

SCENE_SEED = 1234
SOURCE_SEED = 1
TARGET_SEED = 2

# ring count x azimuth steps for the named sizes
SIZES = {
    10_000: (16, 625),
    200_000: (64, 3125),
    500_000: (64, 7813),
    1_000_000: (64, 15625),
}


def rpy_to_matrix(t, rpy_deg):
    """4x4 double transform, R = Rz(yaw) Ry(pitch) Rx(roll)."""
    r, p, y = np.deg2rad(np.asarray(rpy_deg, dtype=np.float64))
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = t
    return T


T_TRUE = rpy_to_matrix((0.20, 0.10, 0.05), (0.5, 0.3, 1.0))


def make_scene(seed: int = SCENE_SEED):
    """Cylinder centres (64,2) and radii (64,), kept at least 3 m from the sensor."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cx = rng.uniform(-BOX_HALF[0] + 2.0, BOX_HALF[0] - 2.0, 4 * N_CYL)
    cy = rng.uniform(-BOX_HALF[1] + 2.0, BOX_HALF[1] - 2.0, 4 * N_CYL)
    rad = rng.uniform(0.3, 1.5, 4 * N_CYL)
    keep = np.hypot(cx, cy) > 3.0 + rad
    idx = np.nonzero(keep)[0][:N_CYL]
    return np.stack([cx[idx], cy[idx]], axis=1), rad[idx]


def _raycast(o, d, scene):
    """Nearest hit distance and world-frame surface normal for rays o + t d (d unit), float64."""
    centres, radii = scene
    n = d.shape[0]
    t_best = np.full(n, np.inf)
    nrm = np.zeros((n, 3))

    def consider(t, normal):
        nonlocal t_best, nrm
        ok = (t > 1e-6) & (t < t_best)
        t_best = np.where(ok, t, t_best)
        nrm[ok] = normal[ok] if normal.ndim == 2 else normal

    with np.errstate(divide="ignore", invalid="ignore"):
        # ground
        t = (GROUND_Z - o[2]) / d[:, 2]
        consider(np.where(np.isfinite(t), t, np.inf), np.array([0.0, 0.0, 1.0]))
        # walls (infinite height; the box is closed laterally)
        for axis, half in ((0, BOX_HALF[0]), (1, BOX_HALF[1])):
            other = 1 - axis
            for sgn in (-1.0, 1.0):
                t = (sgn * half - o[axis]) / d[:, axis]
                hit_other = o[other] + t * d[:, other]
                ok = np.isfinite(t) & (np.abs(hit_other) <= BOX_HALF[other] + 1e-9)
                nvec = np.zeros(3)
                nvec[axis] = -sgn
                consider(np.where(ok, t, np.inf), nvec)
        # cylinders (vertical, infinite height, clipped by the ground test above); only rays whose
        # world azimuth falls inside the cylinder's angular extent seen from o are tested
        phi = np.arctan2(d[:, 1], d[:, 0])
        order = np.argsort(phi, kind="stable")
        phis = phi[order]
        for c, r in zip(centres, radii):
            oc = o[:2] - c
            dist = np.hypot(oc[0], oc[1])
            theta = np.arctan2(-oc[1], -oc[0])
            half = np.arcsin(min(1.0, r / dist)) + 1e-6
            cand = []
            for lo, hi in ((theta - half, theta + half), (theta - half + 2 * np.pi, theta + half + 2 * np.pi),
                           (theta - half - 2 * np.pi, theta + half - 2 * np.pi)):
                a0, a1 = np.searchsorted(phis, lo), np.searchsorted(phis, hi)
                if a1 > a0:
                    cand.append(order[a0:a1])
            if not cand:
                continue
            ci = np.concatenate(cand)
            dxy = d[ci, :2]
            a = np.einsum("ij,ij->i", dxy, dxy)
            b = dxy @ oc
            cc = oc @ oc - r * r
            disc = b * b - a * cc
            sq = np.sqrt(np.where(disc > 0, disc, np.nan))
            t = (-b - sq) / a
            ok = np.isfinite(t) & (t > 1e-6) & (t < t_best[ci])
            ci, t = ci[ok], t[ok]
            t_best[ci] = t
            hit = o[:2] + t[:, None] * d[ci, :2]
            nrm[ci, :2] = (hit - c) / r
            nrm[ci, 2] = 0.0
    return t_best, nrm


def _first_occurrence(pts: np.ndarray) -> np.ndarray:
    """Sorted indices of the first occurrence of every distinct xyz row (bit pattern equality)."""
    b = np.ascontiguousarray(pts).view(np.uint32).astype(np.uint64)
    key = (b[:, 0] * np.uint64(0x9E3779B97F4A7C15)) ^ (b[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F)) \
        ^ (b[:, 2] * np.uint64(0x165667B19E3779F9))
    _, first = np.unique(key, return_index=True)
    first.sort()
    return first


def velodyne_scan(rings: int, az_steps: int, pose=None, noise_seed: int = SOURCE_SEED,
                  scene=None, n_points: int | None = None, sigma: float = SIGMA,
                  return_normals: bool = False):
    """One scan in the sensor frame, fp32 (n,3), azimuth-major point order, de-duplicated.

    pose: 4x4 sensor-to-world transform (identity if None).  n_points: if given, extra azimuth
    steps are appended until exactly n_points remain after dropping misses and duplicates.
    """
    scene = make_scene() if scene is None else scene
    pose = np.eye(4) if pose is None else np.asarray(pose, dtype=np.float64)
    want = rings * az_steps if n_points is None else n_points
    extra = 0
    elev = np.deg2rad(np.linspace(ELEV_DEG[0], ELEV_DEG[1], rings))
    while True:
        steps = az_steps + extra
        az = 2.0 * np.pi * np.arange(steps) / az_steps
        azg, elg = np.meshgrid(az, elev, indexing="ij")  # azimuth-major
        azg, elg = azg.ravel(), elg.ravel()
        d_s = np.stack([np.cos(elg) * np.cos(azg), np.cos(elg) * np.sin(azg), np.sin(elg)], axis=1)
        d_w = d_s @ pose[:3, :3].T
        t, n_w = _raycast(pose[:3, 3], d_w, scene)
        rng_local = np.random.Generator(np.random.PCG64(noise_seed))
        noise = rng_local.normal(0.0, sigma, t.shape[0]) if sigma > 0 else 0.0
        ok = np.isfinite(t) & (t <= MAX_RANGE)
        pts = ((t + noise)[:, None] * d_s)[ok].astype(np.float32)
        nrm = (n_w @ pose[:3, :3])[ok]  # world normal -> sensor frame (R^T n)
        first = _first_occurrence(pts)
        pts, nrm = pts[first], nrm[first]
        if pts.shape[0] >= want or n_points is None:
            break
        extra += max(8, int(1.05 * (want - pts.shape[0]) / rings) + 1)
    pts, nrm = pts[:want], nrm[:want]
    if return_normals:
        # orient towards the sensor (origin of the sensor frame)
        flip = np.einsum("ij,ij->i", nrm, pts.astype(np.float64)) > 0
        nrm[flip] *= -1.0
        return pts, nrm.astype(np.float32)
    return pts


def scan_pair(n: int, scan_id: int | None = None, T_true=None, return_normals: bool = False):
    """(source, target[, target_normals]) of n points each; target = scene seen from T_true^-1.

    scan_id None -> the single-match seeds (source noise 1, target noise 2); batch scan k uses
    noise seeds 1000+k (source) and 2000+k (target) against the same scene.
    """
    rings, az = SIZES[n]
    T_true = T_TRUE if T_true is None else T_true
    s_seed, t_seed = (SOURCE_SEED, TARGET_SEED) if scan_id is None else (1000 + scan_id, 2000 + scan_id)
    scene = make_scene()
    src = velodyne_scan(rings, az, None, s_seed, scene, n_points=n)
    out = velodyne_scan(rings, az, np.linalg.inv(T_true), t_seed, scene, n_points=n,
                        return_normals=return_normals)
    return (src, *out) if return_normals else (src, out)


def scan_batch(n: int, count: int, first_id: int = 0, ids=None):
    """`count` source scans of n points of the same scene from the sensor origin, differing only in
    their range-noise seed (batch scan k: seed 1000 + k), plus the shared target ("map") scan taken
    from T_true^-1 (seed 2).  One ray cast serves all sources, so a 256-scan batch costs seconds."""
    rings, az = SIZES[n]
    scene = make_scene()
    elev = np.deg2rad(np.linspace(ELEV_DEG[0], ELEV_DEG[1], rings))
    azs = 2.0 * np.pi * np.arange(az) / az
    azg, elg = np.meshgrid(azs, elev, indexing="ij")
    azg, elg = azg.ravel(), elg.ravel()
    d_s = np.stack([np.cos(elg) * np.cos(azg), np.cos(elg) * np.sin(azg), np.sin(elg)], axis=1)
    t, _ = _raycast(np.zeros(3), d_s, scene)
    ok = np.isfinite(t) & (t <= MAX_RANGE)
    t, d_s = t[ok], d_s[ok]
    ids = list(range(first_id, first_id + count)) if ids is None else list(ids)
    sources = []
    for k in ids:
        noise = np.random.Generator(np.random.PCG64(1000 + k)).normal(0.0, SIGMA, t.shape[0])
        sources.append(((t + noise)[:, None] * d_s).astype(np.float32))
    target = velodyne_scan(rings, az, np.linalg.inv(T_TRUE), TARGET_SEED, scene, n_points=n)
    return sources, target


def map_cloud(n_scans: int = 5, n_per_scan: int = 1_000_000, baseline: float = 4.0):
    """Union of n_scans scans along a baseline on x, all expressed in the frame of T_true^-1
    (the target frame of scan_pair) - the 5M-point NDT map of BASELINE.json config 4."""
    rings, az = SIZES[n_per_scan]
    scene = make_scene()
    base = np.linalg.inv(T_TRUE)
    clouds = []
    for k in range(n_scans):
        off = np.eye(4)
        off[0, 3] = baseline * (k / max(1, n_scans - 1) - 0.5)
        pose = off @ base
        pts = velodyne_scan(rings, az, pose, 3000 + k, scene, n_points=n_per_scan).astype(np.float64)
        # sensor_k frame -> world -> target frame
        rel = np.linalg.inv(base) @ pose
        clouds.append((pts @ rel[:3, :3].T + rel[:3, 3]).astype(np.float32))
    allp = np.concatenate(clouds, axis=0)
    return allp[_first_occurrence(allp)]


def to_xyzw(pts: np.ndarray) -> np.ndarray:
    """(n,3) fp32 -> (n,4) fp32 in the pcl::PointXYZ memory layout (w = 1.0f)."""
    out = np.ones((pts.shape[0], 4), dtype=np.float32)
    out[:, :3] = pts
    return out
