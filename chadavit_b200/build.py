"""Build libchadavit_b200.so (sm_100a only) in-tree with nvcc.

    python -m chadavit_b200.build [--force] [--verbose]

Each .cu under chadavit_b200/csrc is compiled to an object (in parallel, skipped when up to date) and linked into
chadavit_b200/lib/libchadavit_b200.so with a static cudart.  cuTensorMapEncodeTiled is resolved at run time through
cudaGetDriverEntryPoint, so the library loads on hosts without libcuda (the CPU build/test container).
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
INC = os.path.join(os.path.dirname(ROOT), "include")
# CB_VARIANT=<name> (with CB_NVCC_EXTRA=<flags>) builds a debug variant next to the product library, e.g. the in-kernel
# timeline build: CB_VARIANT=tl CB_NVCC_EXTRA=-DCB_TIMELINE -> lib/libchadavit_b200_tl.so (loaded when CB_VARIANT=tl is set).
VARIANT = os.environ.get("CB_VARIANT", "")
OBJ = os.path.join(ROOT, "csrc", "_obj" + ("_" + VARIANT if VARIANT else ""))
LIBDIR = os.path.join(ROOT, "lib")
LIB = os.path.join(LIBDIR, "libchadavit_b200" + ("_" + VARIANT if VARIANT else "") + ".so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "-I", INC, "-I", CSRC, *os.environ.get("CB_NVCC_EXTRA", "").split()]


def _newest_header() -> float:
    hs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(INC, "*.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src: str, force: bool, verbose: bool) -> str:
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), _newest_header()):
        return obj
    cmd = [NVCC, *FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
