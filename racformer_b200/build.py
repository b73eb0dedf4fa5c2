"""Build libracformer_ops.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Replaces models/csrc/setup.py:95-122 of the reference (which builds a torch extension with no arch flags).
The library has no torch / Python dependency; it links cudart statically so it loads anywhere.

    python -m racformer_b200.build [--force] [--verbose]
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(REPO_ROOT, "include")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.environ.get("RACF_LIB_PATH") or os.path.join(LIB_DIR, "libracformer_ops.so")   # override: tuning experiments
STAMP = LIB_PATH + ".srchash"

SOURCES = ["msmv.cu", "msda.cu", "points.cu", "points_train.cu", "layout.cu", "bev_pool.cu", "mixing.cu", "mixing_bwd.cu", "mixing_bwd_tc.cu", "mixing_tc.cu", "mixing_ws.cu", "linear.cu", "linear_wide.cu", "rowops.cu", "sasa.cu", "sasa_train.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=true",  # FMA contraction on (as in the reference build); no fast-math
    "-Xcompiler", "-fPIC,-O3",
    "-shared", "-cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libracformer_ops.so cannot be built")


COMPILE_FLAGS = [f for f in NVCC_FLAGS if f not in ("-shared", "-cudart", "static")]
HEADERS = lambda: sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [
    os.path.join(INCLUDE, "racformer_ops.h")]


def _digest(paths, extra=""):
    h = hashlib.sha256()
    for f in paths:
        with open(f, "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    h.update((" ".join(NVCC_FLAGS) + os.environ.get("RACF_NVCC_DEFINES", "") + extra).encode())
    return h.hexdigest()


def source_hash():
    return _digest([os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(INCLUDE, "racformer_ops.h")])


def is_current():
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == source_hash()


def build(force=False, verbose=False):
    """Compile the library if sources changed. Returns the path of the .so.

    Each .cu is compiled to its own object (in parallel; an object is reused while its source, the shared headers and
    the flags are unchanged), then linked into one shared library."""
    if not force and is_current():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj" + ("" if LIB_PATH.endswith("libracformer_ops.so") else "_" + os.path.basename(LIB_PATH)))
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    defines = os.environ.get("RACF_NVCC_DEFINES", "").split()
    headers = HEADERS()
    jobs, objs = [], []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(obj_dir, src[:-3] + ".o")
        stamp = obj + ".srchash"
        digest = _digest([path] + headers)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
            continue
        cmd = [nvcc] + COMPILE_FLAGS + defines + ["-I", INCLUDE, "-I", CSRC] + (["-Xptxas", "-v"] if verbose else []) + [
            "-c", path, "-o", obj]
        jobs.append((cmd, stamp, digest, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for cmd, stamp, digest, proc in jobs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            for _c, _s, _d, other in jobs:
                if other.poll() is None:
                    other.kill()
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
        log.append(out)
        with open(stamp, "w") as fh:
            fh.write(digest)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC"] + objs + ["-o", LIB_PATH]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + " ".join(link) + "\n" + res.stdout + res.stderr)
    if verbose:
        print("".join(log) + res.stdout + res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(source_hash())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
