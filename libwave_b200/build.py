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
    return SO


if __name__ == "__main__":
    print(build())
