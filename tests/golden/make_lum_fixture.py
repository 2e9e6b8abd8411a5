#!/usr/bin/env python
"""Golden vectors for the Lu-Milios information estimators.

ICPMatcher::estimateLUM and ::estimateLUMold are first-party reference arithmetic
(wave_matching/src/icp_pcl_functions.cpp:182-289 and :51-179), but they sit behind PCL / Eigen
types, so the reference cannot be compiled in this image.  As for the Censi estimator
(make_censi_fixture.py), this script EXECUTES the reference's own statements: it reads, at
generation time, the two accumulation loops of each function from /root/reference (the M'M / M'Z
sums and the s^2 loop), rewrites them mechanically into Python (Eigen element access `v[ci](k)` ->
`v[ci][k]`, `M(r, c)` -> `M[r, c]`, `pow`, `static_cast<float>`; nothing else) and evaluates them with
numpy scalars, whose float32 / float64 promotion is C++'s for these expressions (float * float ->
float; double op float -> double; `float += float` rounds to float).  Around the loops:
  * the pair vectors are built as the functions build them (estimateLUM :201-214 from the
    correspondence list; estimateLUMold :71-99 from a 1-NN search with the STRICT test
    d2 < max_corr^2, here by brute force in fp32 L2_Simple arithmetic with lowest-index ties);
  * `MM.inverse() * MZ` is Eigen's general 6x6 inverse (PartialPivLU); numpy's LAPACK inverse is the
    same factorisation, equal to rounding (~1e-16 * cond) - which is why D is compared at 1e-9 and
    everything after it through the fp32 `ss` at fp32 resolution;
  * the symmetric fill and `MM * (1.0f / ss)` are restated.
No reference source text is stored in the repository: only inputs and resulting matrices
(tests/golden/lum_fixture.npz).  Run here, where /root/reference exists:

    python tests/golden/make_lum_fixture.py
"""
from __future__ import annotations

import pathlib
import re

import numpy as np

REF = pathlib.Path("/root/reference/wave_matching/src/icp_pcl_functions.cpp")
OUT = pathlib.Path(__file__).resolve().parent / "lum_fixture.npz"


def function_body(name: str) -> str:
    text = REF.read_text()
    start = text.index(f"void ICPMatcher::{name}()")
    nxt = text.find("void ICPMatcher::", start + 10)
    return text[start: nxt if nxt > 0 else len(text)]


def loop_bodies(name: str):
    """(sum loop statements, ss loop statement) of one of the two functions."""
    body = re.sub(r"//[^\n]*", "", function_body(name))
    loops = [m.start() for m in re.finditer(r"for \(int ci = 0; ci != numCorr; \+\+ci\)", body)]
    assert len(loops) == 2, loops

    def block(pos):
        o = body.index("{", pos)
        depth, i = 0, o
        while True:
            depth += body[i] == "{"
            depth -= body[i] == "}"
            if depth == 0:
                return body[o + 1:i]
            i += 1
    return [[" ".join(s.split()) for s in block(p).split(";") if s.strip()] for p in loops]


def to_python(stmt: str) -> str:
    s = re.sub(r"corrs_(aver|diff)\[ci\]\((\d)\)", r"corrs_\1[ci][\2]", stmt)
    s = re.sub(r"\b(MM)\((\d), (\d)\)", r"\1[\2, \3]", s)
    s = re.sub(r"\b(MZ|D)\((\d)\)", r"\1[\2]", s)
    s = s.replace("2.0f", "np.float32(2.0)")
    if s.startswith("ss +="):
        # float ss; ss += static_cast<float>(double expression)
        expr = s[len("ss +="):].strip()
        assert expr.startswith("static_cast<float>(") and expr.endswith(")")
        inner = expr[len("static_cast<float>("):-1]
        return f"ss = np.float32(ss + np.float32({inner}))"
    m = re.match(r"^(MM\[\d, \d\]|MZ\[\d\]) (\+=|-=) (.*)$", s)
    assert m, s
    return f"{m.group(1)} {m.group(2)} {m.group(3)}"


def run_reference(name: str, corrs_aver, corrs_diff):
    """Everything after the pair vectors exist, with the reference's own loop statements."""
    sums, ssl = loop_bodies(name)
    code_sum = compile("\n".join(to_python(s) for s in sums), f"<{name} sums>", "exec")
    code_ss = compile("\n".join(to_python(s) for s in ssl), f"<{name} ss>", "exec")
    numCorr = len(corrs_aver)
    MM, MZ = np.zeros((6, 6)), np.zeros(6)
    env = {"np": np, "pow": lambda a, b: np.float64(a) ** np.float64(b), "MM": MM, "MZ": MZ,
           "corrs_aver": corrs_aver, "corrs_diff": corrs_diff}
    for ci in range(numCorr):
        env["ci"] = ci
        exec(code_sum, env)
    MM[0, 0] = MM[1, 1] = MM[2, 2] = np.float64(np.float32(numCorr))
    for r, c in ((4, 0), (5, 0), (3, 1), (4, 1), (3, 2), (5, 2), (4, 3), (5, 3), (5, 4)):
        MM[r, c] = MM[c, r]
    D = np.linalg.inv(MM) @ MZ
    env["D"] = D
    env["ss"] = np.float32(0.0)
    for ci in range(numCorr):
        env["ci"] = ci
        exec(code_ss, env)
    ss = env["ss"]
    info = MM * np.float64(np.float32(1.0) / ss)
    return MM.copy(), MZ.copy(), D, np.float32(ss), info


def pairs_lum(final, target, q, m):
    """estimateLUM :198-216: 0.5f * (a + b) and a - b in float, for correspondences with index_match > -1."""
    keep = m > -1
    a, b = final[q[keep]].astype(np.float32), target[m[keep]].astype(np.float32)
    aver = (np.float32(0.5) * (a + b)).astype(np.float32)
    diff = (a - b).astype(np.float32)
    return list(aver), list(diff)


def pairs_lum_old(final, target, max_corr):
    """estimateLUMold :71-99: 1-NN of every point of `final` in the target (fp32 L2_Simple, lowest index
    among exact ties), kept when d2 < max_corr^2 (strict, the fp32 distance compared as double)."""
    aver, diff, nn = [], [], []
    t = target.astype(np.float32)
    for p in final.astype(np.float32):
        d = p[None, :] - t
        d2 = ((d[:, 0] * d[:, 0]) + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]        # fp32, left to right
        j = int(np.argmin(d2))                                                     # first minimum = lowest index
        nn.append(j)
        if np.float64(d2[j]) < max_corr * max_corr:
            aver.append((np.float32(0.5) * (p + t[j])).astype(np.float32))
            diff.append((p - t[j]).astype(np.float32))
    return aver, diff, np.asarray(nn, np.int32)


def make_case(seed: int, n: int, t, noise: float):
    rng = np.random.default_rng(seed)
    rg = rng.uniform(2.0, 40.0, n)
    az = rng.uniform(-np.pi, np.pi, n)
    el = np.deg2rad(rng.uniform(-25.0, 3.0, n))
    final = np.stack([rg * np.cos(el) * np.cos(az), rg * np.cos(el) * np.sin(az), rg * np.sin(el)], 1).astype(np.float32)
    target = (final.astype(np.float64) + np.asarray(t) + rng.normal(0, noise, (n, 3))).astype(np.float32)
    target = target[rng.permutation(n)]
    return final, target


def main():
    out = {"n_cases": 3}
    cases = [make_case(21, 400, (0.01, -0.02, 0.005), 0.02), make_case(22, 300, (0.0, 0.0, 0.0), 0.05),
             make_case(23, 500, (0.10, 0.05, -0.02), 0.01)]
    for c, (final, target) in enumerate(cases):
        rng = np.random.default_rng(100 + c)
        max_corr = (0.04, 3.0, 0.118)[c]                    # cases 0 and 2 drop pairs through the strict test
        aver_o, diff_o, nn = pairs_lum_old(final, target, max_corr)
        MMo, MZo, Do, sso, infoo = run_reference("estimateLUMold", aver_o, diff_o)
        # estimateLUM works from icp.correspondences_: a subset of the queries with their matches
        q = np.sort(rng.choice(len(final), size=(3 * len(final)) // 4, replace=False)).astype(np.int32)
        m = nn[q].copy()
        aver, diff = pairs_lum(final, target, q, m)
        MM, MZ, D, ss, info = run_reference("estimateLUM", aver, diff)
        out.update({f"final{c}": final, f"target{c}": target, f"q{c}": q, f"m{c}": m, f"max_corr{c}": max_corr,
                    f"MM{c}": MM, f"MZ{c}": MZ, f"D{c}": D, f"ss{c}": ss, f"info{c}": info,
                    f"old_n{c}": len(aver_o), f"old_MM{c}": MMo, f"old_MZ{c}": MZo, f"old_D{c}": Do, f"old_ss{c}": sso,
                    f"old_info{c}": infoo})
        print(f"case {c}: LUM {len(aver)} pairs ss {ss:.6g} |info| {np.abs(info).max():.4g};  LUMold {len(aver_o)} of "
              f"{len(final)} pairs ss {sso:.6g} |info| {np.abs(infoo).max():.4g}")
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
