"""In-tree build of the sm_100a C-ABI library:  python -m zbot_lab_b200.build

nvcc cross-compiles on a box without a GPU.  The .so lands next to the sources
(``zbot_lab_b200/csrc/libzbot_b200.so``) so it travels with the tree; it is git-ignored.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libzbot_b200.so")
SOURCES = ["zbot_kernels.cu"]


def deps() -> list:
    """Every source the library is built from: all of csrc/*.cu|*.h (globbed, so a new header cannot be forgotten)
    plus the C-ABI header."""
    import glob
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh"))) + [
        os.path.join(HERE, "..", "include", "zbot_b200.h")]


NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(p):
        raise RuntimeError("nvcc not found")
    return p


def is_stale() -> bool:
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in deps())


def build_native(force: bool = False, verbose: bool = False, out: str | None = None, extra: list | None = None) -> str:
    """``out`` / ``extra``: tuning builds (a differently-flagged copy next to the product library, loaded with
    ``ZBOT_B200_LIB=<path>``, see ``native.lib``); the product build uses neither."""
    if out is None and not force and not is_stale():
        return OUT
    out = out or OUT
    cmd = [nvcc_path(), *NVCC_FLAGS, *os.environ.get("ZBOT_NVCC_EXTRA", "").split(), *(extra or []), "-o", out, *SOURCES]
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper without libgomp specs; nvcc only needs a host g++
    if os.access("/usr/bin/g++", os.X_OK):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(CSRC, "build.log" if out == OUT else os.path.basename(out) + ".log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose or r.returncode != 0:
        print(log)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed; see zbot_lab_b200/csrc/build.log")
    return out


if __name__ == "__main__":
    build_native(force="--force" in sys.argv, verbose=True)
    print("built", OUT)
