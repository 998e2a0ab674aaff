"""Builds noise_gnn_b200/libngnn_b200.so (sm_100a only) with nvcc, in-tree.

The library is a plain C-ABI shared object (include/ngnn_b200.h); it has no Python or torch
dependency, so the build is one nvcc invocation per .cu plus a link, cached on source mtimes.
nvcc cross-compiles without a GPU, so this runs in the builder container as well as on the box.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
BUILD_DIR = PKG_DIR / "build"
LIB_PATH = PKG_DIR / "libngnn_b200.so"

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
              "-Xcudafe", "--diag_suppress=177"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libngnn_b200.so cannot be built (set NVCC=/path/to/nvcc)")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps_mtime() -> float:
    files = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(INCLUDE.glob("*.h")) + [Path(__file__)]
    return max(f.stat().st_mtime for f in files)


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources()) or _deps_mtime() > t


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None) -> Path:
    """Compile every csrc/*.cu for sm_100a and link libngnn_b200.so. Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    BUILD_DIR.mkdir(exist_ok=True)
    hdr_t = _deps_mtime()
    flags = ARCH_FLAGS + NVCC_FLAGS + (extra_flags or [])
    jobs = []
    for src in sources():
        obj = BUILD_DIR / (src.stem + ".o")
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_t):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *flags, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for src, r in ex.map(compile_one, jobs):
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
            if verbose and r.stderr.strip():
                print(r.stderr, file=sys.stderr)
    objs = [str(BUILD_DIR / (s.stem + ".o")) for s in sources()]
    tmp = LIB_PATH.with_suffix(".so.tmp")
    cmd = [nvcc, *ARCH_FLAGS, "-shared", "-o", str(tmp), *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
