"""Build libemk.so (sm_100a) in-tree with nvcc.  Used by ``__graft_entry__.build()`` and by
``python -m encodermap_b200._build``.  The shared library lands next to this file so that it
travels with the repository snapshot to the GPU box; it is git-ignored."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
LIB = PKG / "libemk.so"
SOURCES = ["emk_api.cu", "pair_tile.cu", "backmap.cu", "elementwise.cu", "comm.cu", "cart_loss.cu", "generate.cu", "sidechain.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or Path(cand).exists()):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list((PKG.parent / "include").glob("*.h"))
    stamp = BUILD / "stamp"
    digest = _digest(deps)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    nvcc = _nvcc()

    def compile_one(src: str):
        obj = BUILD / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v", "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD / (src + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stderr[-4000:]}")
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC", "-o", str(LIB), *map(str, objs), "-ldl"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    stamp.write_text(digest)
    if verbose:
        for src in SOURCES:
            print((BUILD / (src + ".log")).read_text())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
