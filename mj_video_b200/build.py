"""In-tree build of the CUDA C-ABI library `libmjv_b200.so` for sm_100a.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the tree to the GPU box.
cudart is linked statically and libcuda is resolved at run time, so the library loads (and exports every symbol
of include/mjv_b200.h) on a machine without a driver.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
BUILD_DIR = PKG_DIR / "build"
LIB_PATH = PKG_DIR / "libmjv_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--compiler-options", "-fPIC",
    "-Xptxas", "-v",
    "-I", str(INCLUDE),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stamp(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(src.read_bytes())
    for dep in sorted(list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))):
        h.update(dep.read_bytes())
    return h.hexdigest()


def _compile_one(src: Path) -> tuple[Path, str]:
    obj = BUILD_DIR / (src.stem + ".o")
    stamp_file = BUILD_DIR / (src.stem + ".stamp")
    stamp = _stamp(src)
    if obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return obj, ""
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    (BUILD_DIR / (src.stem + ".ptxas.log")).write_text(res.stderr)
    stamp_file.write_text(stamp)
    return obj, res.stderr


def build(verbose: bool = False) -> Path:
    BUILD_DIR.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    if not sources:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    with cf.ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        results = list(ex.map(_compile_one, sources))
    objs = [str(o) for o, _ in results]
    newest = max(os.path.getmtime(o) for o in objs)
    if not LIB_PATH.exists() or os.path.getmtime(LIB_PATH) < newest:
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
               "-o", str(LIB_PATH), *objs, "-ldl"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    return LIB_PATH


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
