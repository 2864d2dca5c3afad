"""Build recipes for the native parts (in-tree, so the .so files travel with the repo).

  libdwg.so   CUDA kernels (sm_100a) + C ABI + C++ host (loader, compiler, CLI)
  dw          the `dw` command-line front end (thin main over dwg_dw_main)
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.environ.get("DWG_LIB") or os.path.join(PKG, "libdwg.so")  # DWG_LIB: an experimental build (scripts/ only)
DW = os.path.join(PKG, "dw")

CU_SOURCES = ["dwg_api.cu"]
HOST_SOURCES = ["host/compile.cc", "host/loader.cc", "host/cmd_parser.cc", "host/gibbs_app.cc",
                "host/text2bin.cc", "host/synth.cc"]
HEADERS = ["dwg_device.cuh", "dwg_kernels.cuh", "dwg_sell.cuh", "dwg_cell.cuh", "graph_image.h", "host/parallel.h", "host/cmd_parser.h",
           os.path.join(ROOT, "include", "dwg.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _cuda_home():
    return os.path.dirname(os.path.dirname(os.path.realpath(_nvcc())))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """One object per source (the CUDA translation unit takes minutes, the host files seconds: a host
    change does not pay for the kernels again), linked into libdwg.so."""
    from concurrent.futures import ThreadPoolExecutor
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    dev = ["-DDWG_DEV_BUILD"] if os.environ.get("DWG_DEV_BUILD") == "1" else []  # C2-shaped kernels only (see dwg_api.cu)
    dev += os.environ.get("DWG_NVCC_EXTRA", "").split()
    objdir = os.path.join(PKG, "_obj" + ("_" + os.path.basename(LIB) if os.environ.get("DWG_LIB") else ""))
    os.makedirs(objdir, exist_ok=True)
    jobs, objs = [], []
    for src in CU_SOURCES + HOST_SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(objdir, src.replace("/", "_") + ".o")
        objs.append(obj)
        if force or _stale(obj, [path] + hdrs):
            if src.endswith(".cu"):
                jobs.append([_nvcc()] + NVCC_FLAGS + dev + ["-c", "-o", obj, path])
            else:
                jobs.append(["g++", "-O3", "-std=c++17", "-fPIC", "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas",
                             "-I" + os.path.join(_cuda_home(), "include"), "-c", "-o", obj, path])

    def run(cmd):
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    with ThreadPoolExecutor(max_workers=max(1, min(8, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        run([_nvcc(), "-shared", "-o", LIB] + objs + ["-ldl", "-lpthread"])
    main = os.path.join(CSRC, "host", "dw_main.cc")
    if os.path.exists(main) and (force or _stale(DW, [main, LIB])):
        cmd = ["g++", "-O2", "-std=c++17", "-o", DW, main, "-L" + PKG, "-ldwg",
               "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    import sys
    build(force="--force" in sys.argv, verbose=True)
