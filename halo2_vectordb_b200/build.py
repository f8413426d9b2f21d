"""Build the in-tree native libraries with nvcc for sm_100a (no JIT cache: the .so files travel with the tree)."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libh2v.so")
HOSTCHECK = os.path.join(PKG, "libh2v_hostcheck.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources():
    out = [os.path.join(ROOT, "include", "h2v.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h", ".hpp")):
            out.append(os.path.join(CSRC, f))
    return out


def build(force=False, verbose=False):
    srcs = _sources()
    nvcc = os.environ.get("NVCC", "nvcc")
    if force or _newer(LIB, srcs):
        # -split-compile 0: ptxas works on the kernels in parallel (4.5 min -> 2 min on 8 cores), same code
        cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-lineinfo", "-split-compile", "0", *ARCH,
               "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-o", LIB, os.path.join(CSRC, "h2v.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    if force or _newer(HOSTCHECK, srcs):
        cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets",
               "-I" + CSRC, "-o", HOSTCHECK, os.path.join(CSRC, "hostcheck.cu")]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
