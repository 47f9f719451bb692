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
           os.path.join("..", "host", "psk_soft_gpu.hpp"), os.path.join("..", "host", "demo_component.cpp"),
           os.path.join("..", "host", "demo_box.cpp")]
NVCC_COMPILE = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
NVCC_LINK = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared"]


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


def _compile_one(args):
    src, obj, extra, verbose = args
    cmd = [_nvcc()] + NVCC_COMPILE + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res.returncode, res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str | None = None) -> str:
    """Compile every translation unit (in parallel, objects cached under lib/obj) and link libpskd.so.
    `extra_flags` + `out` build a tuning variant (other -D switches) next to the default library."""
    out = out or LIB_PATH
    variant = out != LIB_PATH
    if not force and not variant and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj" + ("_" + os.path.basename(out) if variant else ""))
    os.makedirs(obj_dir, exist_ok=True)
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS[:4])
    hdr_t = max(hdr_t, os.path.getmtime(os.path.abspath(__file__)))
    jobs, objs = [], []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(obj_dir, s[:-3] + ".o")
        objs.append(obj)
        if force or variant or not os.path.isfile(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj, tuple(extra_flags), verbose))
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for src, rc, log in ex.map(_compile_one, jobs):
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + log)
            if verbose:
                print(log)
    cmd = [_nvcc()] + NVCC_LINK + ["-o", out + ".tmp"] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    os.replace(out + ".tmp", out)
    if not variant:
        build_host_demo()
    return out


def build_host_demo() -> str:
    """Compile the C++ host mirror's demo against libpskd.so (checks psk_soft_gpu.hpp builds as plain C++11)."""
    host = os.path.join(HERE, "host")
    exe = os.path.join(LIB_DIR, "demo_component")
    for name in ("demo_component", "demo_box"):      # one component on GPU 0; one bank sharded over every visible GPU
        cmd = ["g++", "-std=c++11", "-O2", "-Wall", "-pthread", "-o", os.path.join(LIB_DIR, name), os.path.join(host, name + ".cpp"),
               "-L" + LIB_DIR, "-lpskd", "-Wl,-rpath,$ORIGIN"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("g++ failed on the host mirror:\n" + res.stdout + res.stderr)
    return exe


if __name__ == "__main__":
    import sys
    # python -m psk_soft_b200._build [--force] [-v] [--variant NAME -DFLAG ...]
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        name, flags = sys.argv[i + 1], [a for a in sys.argv[i + 2:] if a.startswith("-D")]
        print(build(extra_flags=flags, out=os.path.join(LIB_DIR, f"libpskd_{name}.so"), verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
