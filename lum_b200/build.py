"""In-tree build of the native libraries (nvcc, sm_100a only).

    python -m lum_b200.build          # build what is stale
    python -m lum_b200.build --force

The built .so files stay next to this file (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false: one rounding per written operation (bit-parity with the oracle); '/' and sqrtf stay
# IEEE (nvcc defaults -prec-div=true -prec-sqrt=true, no -use_fast_math).
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall",
          "-I", os.path.join(ROOT, "include")]

TARGETS = {
    # the CUDA stage: kernels + C-ABI (include/lumcu.h)
    "liblumcu.so": dict(
        sources=["lumcu.cu"],
        deps=["radiance_device.cuh", "radiance_kernels.cuh", "../../include/lumcu.h"],
        extra=[]),
    # host side above the C-ABI: Lum::Renderer / lum_* mirror (include/lum.hpp, include/clum.h).  Plain C++,
    # built with nvcc as the driver only so one toolchain call covers everything; links liblumcu.so.
    "liblum.so": dict(
        sources=["host/lum_internal.cpp", "host/lum_renderer.cpp", "host/lum_c_api.cpp", "host/vox.cpp"],
        deps=["host/host_math.hpp", "host/vox.hpp", "../../include/lum.hpp", "../../include/lum_opaque.hpp",
              "../../include/lum_math.hpp", "../../include/clum.h", "../../include/lumcu.h", "../liblumcu.so"],
        extra=["-L", HERE, "-llumcu", "-Xlinker", "-rpath,$ORIGIN"]),
}


def _stale(out, files):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(f) > t for f in files if os.path.exists(f)) or os.path.getmtime(__file__) > t


def build_all(force: bool = False, verbose: bool = False) -> list[str]:
    built = []
    for name, spec in TARGETS.items():
        out = os.path.join(HERE, name)
        srcs = [os.path.join(CSRC, s) for s in spec["sources"]]
        deps = srcs + [os.path.normpath(os.path.join(CSRC, d)) for d in spec["deps"]]
        if force or _stale(out, deps):
            cmd = [NVCC, *ARCH, *COMMON, *spec["extra"], "-shared", "-o", out, *srcs]
            if verbose:
                print(" ".join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
            built.append(out)
    return built


if __name__ == "__main__":
    b = build_all(force="--force" in sys.argv, verbose=True)
    print("built:", b if b else "nothing (up to date)")
