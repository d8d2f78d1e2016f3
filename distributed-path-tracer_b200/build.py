"""Builds libptb.so (the C-ABI library with the sm_100a kernels) in-tree.

    python distributed-path-tracer_b200/build.py [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  Flags that matter:
  --fmad=false              closest-hit arithmetic must round like the reference's
                            x86-64 SSE2 build (no fused multiply-add)
  -prec-div/-prec-sqrt      IEEE division and square root (nvcc defaults, stated)
  -Xcompiler -ffp-contract=off   same for the host float code (KD builder, transforms)
  -lineinfo                 so that ncu's source page maps to these files
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
INCLUDE = os.path.join(ROOT, "include")
# A/B builds: PTB_BUILD_TAG=x PTB_NVCC_EXTRA="-DPTB_FOO=1" → libptb_x.so (load it with PTB_LIB=.../libptb_x.so)
TAG = os.environ.get("PTB_BUILD_TAG", "")
OUT = os.path.join(HERE, "libptb" + ("_" + TAG if TAG else "") + ".so")
OBJ = os.path.join(HERE, "build" + ("_" + TAG if TAG else ""))

CU_SOURCES = ["kernels.cu", "extend.cu", "scene.cu", "render.cu", "frame.cu", "api.cu"]
# losing kernel variants kept for A/B measurements (option extend_variant = 0 / 3 / 4): PTB_BUILD_EXPERIMENTS=1
EXPERIMENTS = os.environ.get("PTB_BUILD_EXPERIMENTS", "0") not in ("", "0")
if EXPERIMENTS:
    CU_SOURCES += ["experiments/extend_simple.cu", "experiments/extend_coop.cu", "experiments/extend_ctx.cu"]
CXX_SOURCES = ["kd_build.cpp", "gltf.cpp", "png.cpp"]

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--fmad=false", "-prec-div=true", "-prec-sqrt=true",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-I" + CSRC, "-I" + INCLUDE]
if EXPERIMENTS:
    NVCC_FLAGS.append("-DPTB_BUILD_EXPERIMENTS=1")
NVCC_FLAGS += os.environ.get("PTB_NVCC_EXTRA", "").split()
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-I" + CSRC, "-I" + INCLUDE]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _cxx() -> str:
    # not $CXX: the image exports a wrapper that links libstdc++ statically (see oracle/Makefile)
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _cuda_include() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "include")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd: list[str], verbose: bool) -> None:
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build step failed: " + " ".join(cmd[:3]) + " ...")
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "ptb.h"))
    headers.append(os.path.abspath(__file__))
    objs = []
    nvcc, cxx = _nvcc(), _cxx()
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace("/", "_") + ".o")
        if force or _stale(o, [s] + headers):
            extra = ["-Xptxas", "-v"] if ptxas_info else []
            _run([nvcc, "-ccbin", cxx] + ARCH + NVCC_FLAGS + extra + ["-c", s, "-o", o], verbose or ptxas_info)
        objs.append(o)
    for src in CXX_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        if force or _stale(o, [s] + headers):
            _run([cxx] + CXX_FLAGS + ["-I" + _cuda_include(), "-c", s, "-o", o], verbose)
        objs.append(o)
    if force or _stale(OUT, objs):
        _run([nvcc, "-ccbin", cxx] + ARCH + ["-shared", "-o", OUT] + objs + ["-lz", "-lpthread", "-lrt", "-cudart", "static"],
             verbose)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ptxas_info="--ptxas" in sys.argv)
    print(path)
