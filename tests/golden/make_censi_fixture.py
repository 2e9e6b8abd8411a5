#!/usr/bin/env python
"""Golden vectors for the Censi information estimator.

The estimator's arithmetic is first-party reference code (wave_matching/src/icp.cpp:167-397) but it
sits behind PCL types, so the reference cannot be compiled in this image.  This script instead
EXECUTES the reference's own statements: it reads the body of the correspondence loop of
ICPMatcher::estimateCensi from /root/reference at generation time, rewrites the C++ statements
mechanically into Python (indexing, std:: functions; nothing else) and evaluates them with numpy
scalars, whose float32/float64 promotion rules are the ones C++ applies to these expressions
(float * float -> float, int literal * float -> float, float * double -> double).  The Eigen pieces
around the loop - eulerAngles(0, 1, 2) of the result's rotation, the diagonal measurement
covariance, j * S * j^T, d2J_dZdX * cov_Z * d2J_dZdX^T, the two inverses - are restated from Eigen's
documented algorithms below.

No reference source text is stored in the repository: only the inputs and the resulting matrices
(tests/golden/censi_fixture.npz).  Run here, where /root/reference exists:

    python tests/golden/make_censi_fixture.py
"""
from __future__ import annotations

import pathlib
import re

import numpy as np

REF = pathlib.Path("/root/reference/wave_matching/src/icp.cpp")
OUT = pathlib.Path(__file__).resolve().parent / "censi_fixture.npz"


def loop_statements():
    """The statements between the `const float &Z1 = ...` declaration and `middle.noalias() += ...`."""
    text = REF.read_text()
    start = text.index("rg = std::sqrt(Z1 * Z1")
    end = text.index("middle.noalias()")
    body = re.sub(r"//[^\n]*", "", text[start:end])
    stmts = [" ".join(s.split()) for s in body.split(";")]
    return [s for s in stmts if s]


def to_python(stmt: str) -> str:
    if stmt.startswith("cov_Z ="):
        return "cov_Z = j @ S @ j.T"
    m = re.match(r"^(\w+)\((\d+), (\d+)\) (\+?=) (.*)$", stmt)
    if m:
        name, i, k, op, expr = m.groups()
        lhs = f"{name}[{i}, {k}]"
    else:
        m = re.match(r"^(rg|br|az) = (.*)$", stmt)
        if not m:
            raise ValueError("unexpected statement: " + stmt)
        lhs, op, expr = m.group(1), "=", m.group(2)
        expr = f"np.float64({expr})"       # the three are declared double
    expr = expr.replace("std::sqrt", "np.sqrt").replace("std::atan2", "np.arctan2").replace("std::atan", "np.arctan")
    expr = re.sub(r"\bcos\(", "np.cos(", expr)
    expr = re.sub(r"\bsin\(", "np.sin(", expr)
    return f"{lhs} {op} {expr}"


def euler_angles_012(R: np.ndarray) -> np.ndarray:
    """Eigen 3.3 MatrixBase::eulerAngles(0, 1, 2) (Geometry/EulerAngles.h), Graphics-Gems style with
    the first angle folded into [0, pi]."""
    i, j, k, odd = 0, 1, 2, False
    res = np.zeros(3)
    res[0] = np.arctan2(R[j, k], R[k, k])
    c2 = np.hypot(R[i, i], R[i, j])
    if (odd and res[0] < 0) or ((not odd) and res[0] > 0):
        res[0] = res[0] - np.pi if res[0] > 0 else res[0] + np.pi
        res[1] = np.arctan2(-R[i, k], -c2)
    else:
        res[1] = np.arctan2(-R[i, k], c2)
    s1, c1 = np.sin(res[0]), np.cos(res[0])
    res[2] = np.arctan2(s1 * R[k, i] - c1 * R[j, i], c1 * R[j, j] - s1 * R[k, j])
    return -res if not odd else res


def polar_rotation(L: np.ndarray) -> np.ndarray:
    """Eigen Transform::rotation() for an Affine transform: U diag(1, 1, sign det) V^T of the linear part."""
    U, _, Vt = np.linalg.svd(L)
    D = np.diag([1.0, 1.0, np.sign(np.linalg.det(U @ Vt))])
    return U @ D @ Vt


def censi_reference(target: np.ndarray, ref: np.ndarray, corr_q, corr_m, T: np.ndarray, lin: float, ang: float):
    """ICPMatcher::estimateCensi evaluated with the reference's own loop body."""
    code = compile("\n".join(to_python(s) for s in loop_statements()), "<estimateCensi loop>", "exec")
    e = euler_angles_012(polar_rotation(T[:3, :3]))
    env = {"np": np, "pow": pow}
    env.update(X1=np.float64(T[0, 3]), X2=np.float64(T[1, 3]), X3=np.float64(T[2, 3]),
               cr=np.cos(e[0]), sr=np.sin(e[0]), cp=np.cos(e[1]), sp=np.sin(e[1]), cy=np.cos(e[2]), sy=np.sin(e[2]))
    env["S"] = np.diag([lin, ang, ang, lin, ang, ang]).astype(np.float64)
    env["j"] = np.zeros((6, 6))
    env["cov_Z"] = np.zeros((6, 6))
    env["d2J_dX2"] = np.zeros((6, 6))
    D = np.zeros((6, 6))
    D[3, 0] = D[4, 1] = D[5, 2] = -2
    env["d2J_dZdX"] = D
    middle = np.zeros((6, 6))
    for q, m in zip(corr_q, corr_m):
        z = np.concatenate([target[m, :3], ref[q, :3]]).astype(np.float32)
        for n, v in zip(("Z1", "Z2", "Z3", "Z4", "Z5", "Z6"), z):
            env[n] = np.float32(v)
        exec(code, env)
        middle += env["d2J_dZdX"] @ env["cov_Z"] @ env["d2J_dZdX"].T
    H = env["d2J_dX2"]
    Hs = np.triu(H) + np.triu(H, 1).T
    inv = np.linalg.inv(Hs)
    info = np.linalg.inv(inv @ middle @ inv)
    return Hs, middle, info, e


def make_case(seed: int, n: int, rpy_deg, t):
    rng = np.random.default_rng(seed)
    # lidar-like returns: ranges 2-60 m, elevations -25..+3 deg
    rg = rng.uniform(2.0, 60.0, n)
    az = rng.uniform(-np.pi, np.pi, n)
    el = np.deg2rad(rng.uniform(-25.0, 3.0, n))
    ref = np.stack([rg * np.cos(el) * np.cos(az), rg * np.cos(el) * np.sin(az), rg * np.sin(el)], 1).astype(np.float32)
    r, p, y = np.deg2rad(rpy_deg)
    Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = t
    T = T.astype(np.float32).astype(np.float64)          # result = final_transformation.cast<double>()
    target = (ref @ T[:3, :3].T + T[:3, 3] + rng.normal(0, 0.02, (n, 3))).astype(np.float32)
    target = target[rng.permutation(n)]
    # correspondences: a subset of the queries, each matched to some target index
    keep = np.sort(rng.choice(n, size=(3 * n) // 4, replace=False)).astype(np.int32)
    match = rng.integers(0, n, size=keep.size).astype(np.int32)
    return ref, target, keep, match, T


def main():
    lin, ang = 2.5e-4, 7.78e-9                              # icp.hpp defaults
    cases = [make_case(11, 160, (0.5, 0.3, 1.0), (0.20, 0.10, 0.05)),
             make_case(12, 200, (-0.7, 0.4, -2.0), (-0.30, 0.05, 0.02)),   # negative roll: folded Euler angles
             make_case(13, 120, (3.0, -2.0, 15.0), (1.0, -0.5, 0.1))]
    out = {"lin": lin, "ang": ang, "n_cases": len(cases)}
    for c, (ref, target, q, m, T) in enumerate(cases):
        H, middle, info, e = censi_reference(target, ref, q, m, T, lin, ang)
        out.update({f"ref{c}": ref, f"target{c}": target, f"q{c}": q, f"m{c}": m, f"T{c}": T,
                    f"H{c}": H, f"middle{c}": middle, f"info{c}": info, f"euler{c}": e})
        print(f"case {c}: {len(q)} pairs, euler {e}, |H| {np.abs(H).max():.3e}, |middle| {np.abs(middle).max():.3e}, "
              f"|info| {np.abs(info).max():.3e}")
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
