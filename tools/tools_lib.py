"""Build and load tools/lib/libracformer_tools.so (measurement aids, tools/csrc/). Not part of the product library."""
import ctypes
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "lib", "libracformer_tools.so")
SOURCES = [os.path.join(HERE, "csrc", "ceiling.cu")]


def build(force=False):
    newest = max(os.path.getmtime(s) for s in SOURCES + [os.path.join(HERE, "csrc", "racformer_tools.h")])
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "-shared", "-cudart", "static", "-I", os.path.join(HERE, "csrc")] + SOURCES + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB


def load():
    lib = ctypes.CDLL(build())
    ll, vp, i = ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int
    lib.racf_bench_gather_ceiling.restype = i
    lib.racf_bench_gather_ceiling.argtypes = [vp, ll, ll, i, vp, vp]
    lib.racf_bench_scatter_ceiling.restype = i
    lib.racf_bench_scatter_ceiling.argtypes = [vp, ll, ll, vp]
    return lib


if __name__ == "__main__":
    print(build(force=True))
