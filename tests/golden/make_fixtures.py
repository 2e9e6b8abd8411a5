"""Regenerates tests/golden/testscan_xyz.f32 from the reference's only real-data fixture.

Source: /root/reference/wave_matching/tests/data/testscan.pcd (PCD v0.7, DATA binary, 55 067
records of 32 bytes, x/y/z fp32 at offsets 0/4/8).  The reference tests
load it with pcl::io::loadPCDFile into a PointCloud<PointXYZ> (tests/icp_tests.cpp:26), i.e. only
x, y, z survive - which is exactly what this script keeps (raw little-endian fp32, n x 3).
Run it in the build container only; /root/reference does not exist on the GPU box.
"""
import hashlib
import pathlib
import sys

import numpy as np

SRC = pathlib.Path("/root/reference/wave_matching/tests/data/testscan.pcd")
DST = pathlib.Path(__file__).with_name("testscan_xyz.f32")


def read_pcd_xyz(path):
    raw = path.read_bytes()
    pos, header = 0, {}
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii").strip()
        pos = end + 1
        if line.startswith("#") or not line:
            continue
        key, *vals = line.split()
        header[key] = vals
        if key == "DATA":
            break
    assert header["DATA"] == ["binary"]
    sizes = [int(s) * int(c) for s, c in zip(header["SIZE"], header["COUNT"])]
    stride, n = sum(sizes), int(header["POINTS"][0])
    offs = dict(zip(header["FIELDS"], np.cumsum([0] + sizes[:-1])))
    body = np.frombuffer(raw, dtype=np.uint8, count=n * stride, offset=pos).reshape(n, stride)
    cols = [body[:, offs[f]:offs[f] + 4].copy().view("<f4")[:, 0] for f in ("x", "y", "z")]
    return np.stack(cols, axis=1)


if __name__ == "__main__":
    xyz = read_pcd_xyz(SRC)
    assert xyz.shape == (55067, 3) and np.isfinite(xyz).all()
    DST.write_bytes(xyz.astype("<f4").tobytes())
    print(DST, xyz.shape, hashlib.sha256(DST.read_bytes()).hexdigest()[:16], file=sys.stderr)
