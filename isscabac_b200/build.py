"""Builds isscabac_b200/libisscabac.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libisscabac.so")
SOURCES = ["kernels.cu", "kernels_lat.cu", "symbols.cu", "iss_stats.cu", "stats.cu", "quantize.cu", "multi_gpu.cu", "host_api.cu", "handle.cu", "dispatch.cpp", "container.cpp"]
HEADERS = ["cabac_lane.cuh", "cabac_wide.cuh", "cabac_spec.cuh", "wide_common.cuh", "bin_emit.cuh", "internal.h", "codec_params.h", os.path.join("..", "..", "include", "isscabac.h"),
           os.path.join("..", "..", "include", "SimpleCABAC.hpp")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--shared", "-cudart", "shared", "--threads", "0", "-ldl"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.exists(c) or c == "nvcc"):
            return c
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS if os.path.exists(os.path.join(CSRC, s))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc()] + NVCC_FLAGS + ["-x", "cu"] + srcs + ["-o", LIB]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libisscabac.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


def build_variant(name: str, defines: list[str]) -> str:
    """Tuning aid (tools/tune.py): the same sources with extra -D flags -> isscabac_b200/_variants/lib<name>.so.
    Select it at run time with ISSCABAC_LIB=<path> (see _lib.py)."""
    out_dir = os.path.join(HERE, "_variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"lib{name}.so")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-x", "cu"] + srcs + ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed building variant {name}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
