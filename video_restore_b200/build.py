"""In-tree build of libvrb200.so (the C-ABI shared library) with nvcc. (The oracle is pure Python / numpy / torch -- the
reference is Python -- so there is nothing else to compile; `--prof` builds an instrumented second library.)

`python -m video_restore_b200.build` or `__graft_entry__.build()`. Cross-compiles for sm_100a without a GPU.
Objects are rebuilt only when a source or header is newer than the object.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = PKG / "csrc" / "_obj"
LIB = PKG / "libvrb200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", str(ROOT / "include"),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libvrb200.so cannot be built (there is no CPU fallback)")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(verbose: bool = False, force: bool = False, extra_flags=(), lib: Path = LIB, obj_dir: Path = OBJ) -> Path:
    """extra_flags / lib / obj_dir: an instrumented second library next to the product one, e.g.
    `python -m video_restore_b200.build --prof` -> libvrb200_prof.so with -DVR_K4_PROF (select it with VR_LIB=...)."""
    nvcc = _nvcc()
    OBJ_ = obj_dir
    OBJ_.mkdir(parents=True, exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "vrb200.h"]
    jobs = []
    for src in sources:
        obj = OBJ_ / (src.stem + ".o")
        if force or _stale(obj, [src, *headers]):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    objs = [OBJ_ / (s.stem + ".o") for s in sources]
    if force or jobs or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-o", str(lib), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-Xlinker", "--no-undefined"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    if "--prof" in sys.argv:
        p = build(verbose="-v" in sys.argv, force="-f" in sys.argv, extra_flags=("-DVR_K4_PROF",),
                  lib=PKG / "libvrb200_prof.so", obj_dir=PKG / "csrc" / "_obj_prof")
    else:
        p = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(p)
