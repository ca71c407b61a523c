"""Builds libvbn_cuda.so in-tree with nvcc for sm_100a.  `python -m vectorizedbayesiannetwork_b200.build`"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libvbn_cuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = os.path.join(PKG_DIR, "build")


def sources():
    """One translation unit per kernel family (compiled in parallel): the C ABI + reduction / KDE kernels, the
    FP32-pipe schedule kernels (heavy, light 4-row, light 1-2-row shapes) and the tcgen05 schedule kernels."""
    return [os.path.join(CSRC, f) for f in ("vbn_cuda.cu", "vbn_k_heavy.cu", "vbn_k_light4.cu", "vbn_k_light2.cu",
                                            "vbn_k_tc.cu", "vbn_k_kde_tc.cu")]


def _deps():
    out = [os.path.join(ROOT, "include", "vbn_cuda.h")]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def _obj_stale(src: str, obj: str) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src, os.path.join(ROOT, "include", "vbn_cuda.h")]
    name = os.path.basename(src)
    for f in os.listdir(CSRC):
        if not f.endswith((".cuh", ".h")):
            continue
        if f == "vbn_schedule_tc.cuh" and name not in ("vbn_k_tc.cu", "vbn_k_kde_tc.cu"):
            continue  # only the tcgen05 unit includes it
        if f in ("vbn_kde.cuh", "vbn_reduce.cuh") and name != "vbn_cuda.cu":
            continue
        deps.append(os.path.join(CSRC, f))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OBJ_DIR, exist_ok=True)
    base = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        base += ["-Xptxas", "-v"]
    base += os.environ.get("VBN_NVCC_EXTRA", "").split()  # dev experiments (-D switches)

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        if force or _obj_stale(src, obj):
            proc = subprocess.run(base + ["-c", src, "-o", obj], capture_output=True, text=True)
            if proc.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + proc.stdout + proc.stderr)
            if verbose:
                sys.stderr.write(proc.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(sources())) as pool:
        objs = list(pool.map(compile_one, sources()))
    proc = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", LIB_PATH],
                          capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
