"""Builds libwave_b200/libwavecu.so (sm_100a) in-tree with nvcc; nvcc cross-compiles without a GPU."""
from __future__ import annotations

import pathlib
import subprocess

_DIR = pathlib.Path(__file__).resolve().parent
SO = _DIR / "libwavecu.so"


def build(force: bool = False) -> pathlib.Path:
    csrc = _DIR / "csrc"
    srcs = list(csrc.glob("*.cu")) + list(csrc.glob("*.cuh")) + [csrc / "Makefile",
                                                                 _DIR.parent / "include" / "wavecu.h"]
    stale = (not SO.exists()) or any(s.stat().st_mtime > SO.stat().st_mtime for s in srcs)
    if force or stale:
        r = subprocess.run(["make", "-C", str(csrc), "-j8"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc build of libwavecu.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    build_shim(force)
    return SO


def build_shim(force: bool = False) -> pathlib.Path:
    """libwave_matching.so (C++ classes with the reference's interface) + the re-stated gtest cases."""
    host = _DIR.parent / "src" / "host"
    args = ["make", "-C", str(host), "-j4"] + (["-B"] if force else [])
    r = subprocess.run(args, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ build of the host shim failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return _DIR / "libwave_matching.so"


if __name__ == "__main__":
    print(build())
