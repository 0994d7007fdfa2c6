"""In-tree build of the CUDA library and the C host tools (nvcc / gcc, no torch)."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "librtlsdr_gpu_scan.so")
HOST_DIR = os.path.join(ROOT, "host")
HOST_BUILD = os.path.join(HOST_DIR, "_build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [
        os.path.join(ROOT, "include", "rtlsdr_gpu_scan.h")]


def build_cuda(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -> rtlsdr_b200/librtlsdr_gpu_scan.so"""
    if not force and _newer(LIB, cuda_sources()):
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB, os.path.join(CSRC, "scan_abi.cu")]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    subprocess.run(cmd, check=True)
    return LIB


def build_host(force=False):
    """gcc: the rtl_power-compatible host program and its library (host/Makefile)."""
    if os.path.exists(os.path.join(HOST_DIR, "Makefile")):
        subprocess.run(["make", "-C", HOST_DIR] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return HOST_BUILD
