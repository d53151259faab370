"""In-tree builds of the native pieces (nvcc cross-compiles sm_100a without a GPU).

  libmapquik_b200.so  CUDA kernels + C ABI (include/mapquik_b200.h)   csrc/mq_lib.cu, csrc/mq_pack.cpp
  libmq_host.so       host-only helper: genome / read simulator        csrc/mq_sim.cpp
"""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libmapquik_b200.so")
HOSTLIB = os.path.join(PKG, "libmq_host.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, "mq_lib.cu"), os.path.join(CSRC, "mq_pack.cpp"), os.path.join(CSRC, "mq_kernels.cuh"),
            os.path.join(CSRC, "mq_scan.cuh"), os.path.join(PKG, "..", "include", "mapquik_b200.h")]
    if force or _stale(LIB, srcs):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, srcs[0], srcs[1]]
        subprocess.check_call(cmd)
    return LIB


def build_host(force=False):
    srcs = [os.path.join(CSRC, "mq_sim.cpp")]
    if force or _stale(HOSTLIB, srcs):
        subprocess.check_call(["g++", "-O3", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-o", HOSTLIB] + srcs)
    return HOSTLIB


def build_all(force=False):
    return build_cuda(force), build_host(force)
