"""Builds ``libgandanet_sm100.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: ``python -m gan_danet_b200.build`` or ``build_library()``.  Objects are cached by source mtime.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libgandanet_sm100.so")
SOURCES = ["core.cu", "igemm_simt.cu", "elementwise.cu", "resample.cu", "attention.cu", "losses.cu", "pam_tc.cu", "conv_tc.cu", "linear_tc.cu", "thin_conv.cu", "ensemble.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(verbose: bool = False, force: bool = False) -> str:
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc_common.cuh"), os.path.join(HERE, "..", "include", "gandanet.h")]

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            with open(obj + ".log", "w") as f:
                f.write(r.stdout + r.stderr)
            if verbose:
                print(f"[build] {src}\n{r.stderr}", file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    try:
        build_harness()
    except Exception as e:      # the harness is a convenience binary: the library (the product) is already built
        print(f"[build] cabi_bench not built: {e}", file=sys.stderr)
    return LIB_PATH


def build_harness() -> str:
    """tools/cabi_bench.cpp: the Python-free harness over the C ABI (plain g++, links the library, rpath = its own directory)."""
    src = os.path.join(HERE, "..", "tools", "cabi_bench.cpp")
    exe = os.path.join(HERE, "cabi_bench")
    if os.path.exists(src) and _stale(exe, [src, LIB_PATH, os.path.join(HERE, "..", "include", "gandanet.h")]):
        cuda = os.path.dirname(os.path.dirname(os.path.realpath(_nvcc()))) if os.path.isabs(_nvcc()) else "/usr/local/cuda"
        cmd = ["g++", "-O2", "-std=c++17", src, "-I" + os.path.join(HERE, "..", "include"), "-I" + os.path.join(cuda, "include"), "-L" + HERE, "-lgandanet_sm100",
               "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath,$ORIGIN", "-o", exe]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"cabi_bench build failed:\n{r.stdout}\n{r.stderr}")
    return exe


if __name__ == "__main__":
    print(build_library(verbose="-v" in sys.argv, force="-f" in sys.argv))
