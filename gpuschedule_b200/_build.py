"""In-tree build of libgsched.so for sm_100a (explicit nvcc; no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "gsched.cu")
SRC_HORUS = os.path.join(HERE, "csrc", "gs_horus.cu")          # utilisation-aware placement engine (gsched_horus.h)
SRC_LOGCOL = os.path.join(HERE, "csrc", "gs_logcol.cpp")       # host side of the log writer (sampled cluster.csv column)
OUT = os.path.join(HERE, "libgsched.so")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-I", os.path.join(REPO, "include")]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build(force=False, verbose=False):
    csrc = os.path.dirname(SRC)
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".cpp", ".h"))]
    deps += [os.path.join(REPO, "include", "gsched.h"), os.path.join(REPO, "include", "gsched_horus.h")]
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps)):
        return OUT
    cmd = [nvcc_path()] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC, SRC_HORUS, SRC_LOGCOL]
    subprocess.run(cmd, check=True, cwd=REPO)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
