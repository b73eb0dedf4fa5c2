"""Recipe: compile the reference's UNMODIFIED MSMV CUDA extension from where its sources lie under /root/reference
into oracle/_ref/ (git-ignored, travels to the GPU box). TEST INFRASTRUCTURE: used only by
tests/test_gpu_parity.py::test_msmv_vs_reference_cuda_extension and by bench.py's optional `ref_cuda` report.

    python -m oracle.build_ref            # ~6 min (nvcc compiles the reference's three translation units)

No reference source is copied: torch.utils.cpp_extension is pointed at
/root/reference/models/csrc/msmv_sampling/{msmv_sampling.cpp,msmv_sampling_forward.cu,msmv_sampling_backward.cu}
and only the build products land in oracle/_ref/. The mmcv MSDA kernels are third-party (mmcv-full==1.6.0), are
not under /root/reference and cannot be built here.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SO = os.path.join(OUT, "_msmv_sampling_cuda.so")
REF = os.environ.get("RACFORMER_REFERENCE", "/root/reference")
SRC_DIR = os.path.join(REF, "models", "csrc", "msmv_sampling")


def build(verbose=False):
    if os.path.exists(SO):
        return SO
    if not os.path.isdir(SRC_DIR):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils import cpp_extension
    cpp_extension.load(
        name="_msmv_sampling_cuda",
        sources=[os.path.join(SRC_DIR, f) for f in
                 ("msmv_sampling.cpp", "msmv_sampling_forward.cu", "msmv_sampling_backward.cu")],
        extra_include_paths=[SRC_DIR], build_directory=OUT, is_python_module=False, verbose=verbose)
    return SO if os.path.exists(SO) else None


POOL_SO = os.path.join(OUT, "bev_pool_v2_ext.so")
POOL_SRC = os.path.join(REF, "models", "csrc", "bev_pool_v2", "src")


def build_bev_pool(verbose=False):
    """The reference's BEVPoolv2 extension (models/csrc/bev_pool_v2/src/{bev_pool.cpp,bev_pool_cuda.cu}), unmodified."""
    if os.path.exists(POOL_SO):
        return POOL_SO
    if not os.path.isdir(POOL_SRC):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils import cpp_extension
    cpp_extension.load(name="bev_pool_v2_ext",
                       sources=[os.path.join(POOL_SRC, f) for f in ("bev_pool.cpp", "bev_pool_cuda.cu")],
                       build_directory=OUT, is_python_module=False, verbose=verbose)
    return POOL_SO if os.path.exists(POOL_SO) else None


def load_prebuilt_bev_pool():
    if not os.path.exists(POOL_SO):
        return None
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location("bev_pool_v2_ext", POOL_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_prebuilt():
    """Import oracle/_ref/_msmv_sampling_cuda.so if it exists (needs torch + a CUDA device to be useful)."""
    if not os.path.exists(SO):
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location("_msmv_sampling_cuda", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="--verbose" in sys.argv))
    print(build_bev_pool(verbose="--verbose" in sys.argv))
