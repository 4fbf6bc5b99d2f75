"""Compile the CUDA sources in ``csrc/`` into the in-tree C-ABI library ``libhp_b200.so``.

sm_100a only (``-gencode arch=compute_100a,code=sm_100a``), ``-lineinfo`` so ncu's source page maps
back to the kernels.  Invoked by ``__graft_entry__.build()``; never at import time.
"""

from __future__ import annotations

import hashlib
import os
import pathlib
import shutil
import subprocess
import sys

PKG = pathlib.Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libhp_b200.so"
STAMP = PKG / "csrc" / ".build_stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--shared", "-cudart", "shared",
]  # fmt: skip


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _fingerprint(files) -> str:
    h = hashlib.sha256()
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def sources():
    return sorted(CSRC.glob("*.cu"))


def build(force: bool = False, verbose: bool = False) -> pathlib.Path:
    srcs = sources()
    deps = srcs + sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*.h"))
    fp = _fingerprint(deps)
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text() == fp:
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(ROOT / "include"), "-I", str(CSRC)]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [str(s) for s in srcs] + ["-o", str(LIB), "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libhp_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    STAMP.write_text(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
