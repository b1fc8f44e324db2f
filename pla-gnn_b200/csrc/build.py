"""Builds libplagnn.so (sm_100a only) in-tree with explicit nvcc calls.

    python pla-gnn_b200/csrc/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is written next to the package's Python files so
that it travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OUT = os.path.join(PKG, "libplagnn.so")
OBJ_DIR = os.path.join(HERE, "_obj")
SOURCES = ["graph_build.cu", "spmm.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm_tma.cu", "gemm_narrow.cu", "gemm_api.cu", "elementwise.cu", "scoring.cu", "preprocess.cu", "gnn32_engine.cu", "dist_comm.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--use_fast_math=false" if False else "-Xcompiler", "-Wall",
         "-I", os.path.join(ROOT, "include")]


def _nccl_include() -> str:
    """nccl.h of the NCCL that PyTorch bundles (types only: the library is resolved with dlopen at run time)."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for loc in (spec.submodule_search_locations or []) if spec else []:
            inc = os.path.join(loc, "include")
            if os.path.exists(os.path.join(inc, "nccl.h")):
                return inc
    except Exception:
        pass
    return "/usr/include"


FLAGS += ["-I", _nccl_include()]
DIAG = "--diag" in sys.argv or os.environ.get("PLAGNN_TMA_DIAG") == "1"
if DIAG:     # diagnostics build of the TMA GEMM (tools/gemm_trace.py, PLAGNN_TMA_DEBUG) -> libplagnn_diag.so, own objects
    FLAGS += ["-DPLAGNN_TMA_DIAG=1"]
    OUT = os.path.join(PKG, "libplagnn_diag.so")
    OBJ_DIR = os.path.join(HERE, "_obj_diag")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(HERE, "common.cuh"), os.path.join(ROOT, "include", "plagnn.h"), os.path.abspath(__file__)]
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr + "\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {cmd[-3]}")
    if force or jobs or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
