"""CPU test: the PCL stand-ins that let the reference's tests build literally - a binary-PCD reader
behind pcl::io::loadPCDFile (SURVEY.md Appendix B layout), pcl::transformPointCloud, and the
implicit constructors of the reference's parameter structs.  The PCD files are written here from
the committed fixture with the header the reference's tests/data/testscan.pcd carries."""
import pathlib
import subprocess

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]


def write_binary_pcd(path, xyz):
    """FIELDS x y z _ intensity ring _ / SIZE 4 4 4 1 4 2 1 / COUNT 1 1 1 4 1 1 10: 32-byte records."""
    n = xyz.shape[0]
    header = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z _ intensity ring _\n"
              "SIZE 4 4 4 1 4 2 1\nTYPE F F F U F U U\nCOUNT 1 1 1 4 1 1 10\n"
              f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA binary\n")
    rec = np.zeros((n, 32), dtype=np.uint8)
    rec[:, 0:12] = np.ascontiguousarray(xyz, dtype="<f4").view(np.uint8).reshape(n, 12)
    rec[:, 16:20] = np.full(n, 100.0, dtype="<f4").view(np.uint8).reshape(n, 4)
    rec[:, 20:22] = (np.arange(n) % 32).astype("<u2").view(np.uint8).reshape(n, 2)
    path.write_bytes(header.encode("ascii") + rec.tobytes() + b"\0" * 100)   # trailing bytes are ignored


def write_ascii_pcd(path, xyz):
    n = xyz.shape[0]
    lines = ["# .PCD v0.7", "VERSION 0.7", "FIELDS x y z intensity", "SIZE 4 4 4 4", "TYPE F F F F", "COUNT 1 1 1 1",
             f"WIDTH {n}", "HEIGHT 1", "VIEWPOINT 0 0 0 1 0 0 0", f"POINTS {n}", "DATA ascii"]
    lines += [f"{float(x)!r} {float(y)!r} {float(z)!r} 7" for x, y, z in xyz.astype(np.float64)]
    path.write_text("\n".join(lines) + "\n")


def test_pcd_reader_transform_and_implicit_ctors(tmp_path, testscan):
    exe = tmp_path / "test_pcd"
    inc = ROOT / "include"
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", f"-I{inc}", f"-I{inc / 'wave/matching/compat'}",
                        "-o", str(exe), str(ROOT / "tests" / "cpp" / "test_pcd.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    write_binary_pcd(tmp_path / "scan.pcd", testscan)
    write_ascii_pcd(tmp_path / "head.pcd", testscan[:100])
    r = subprocess.run([str(exe), str(tmp_path / "scan.pcd"), str(tmp_path / "head.pcd"),
                        str(ROOT / "tests" / "golden" / "testscan_xyz.f32")], capture_output=True, text=True)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout + r.stderr
