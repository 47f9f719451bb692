"""Build recipe for libpskd.so (hand-written sm_100a CUDA + the C-ABI host runtime).

Plain ``nvcc`` -- no torch extension machinery: the library has no torch types at its boundary
(include/pskd.h) and links only cudart.  The .so is built IN-TREE (psk_soft_b200/lib/) so that
it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libpskd.so")
SOURCES = ["pskd_api.cu", "pskd_kernels.cu", "pskd_fused.cu", "pskd_synth.cu"]
HEADERS = ["pskd_exact.cuh", "pskd_internal.h", "pskd_device.cuh", os.path.join("..", "..", "include", "pskd.h"),
           os.path.join("..", "host", "psk_soft_gpu.hpp"), os.path.join("..", "host", "demo_component.cpp")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: libpskd.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.isfile(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH + ".tmp"] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    build_host_demo()
    return LIB_PATH


def build_host_demo() -> str:
    """Compile the C++ host mirror's demo against libpskd.so (checks psk_soft_gpu.hpp builds as plain C++11)."""
    host = os.path.join(HERE, "host")
    exe = os.path.join(LIB_DIR, "demo_component")
    cmd = ["g++", "-std=c++11", "-O2", "-Wall", "-pthread", "-o", exe, os.path.join(host, "demo_component.cpp"),
           "-L" + LIB_DIR, "-lpskd", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed on the host mirror:\n" + res.stdout + res.stderr)
    return exe


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
