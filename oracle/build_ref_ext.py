"""Compile the REFERENCE's own shift_cuda extension into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

Sources are taken where they lie under /root/reference/model/Temporal_shift/cuda/ (shift_cuda.cpp,
shift_cuda_kernel.cu); nothing is copied into this repository.  torch 2.11 no longer accepts
``AT_DISPATCH_FLOATING_TYPES(input.type(), ...)`` (5 sites, shift_cuda_kernel.cu:413,450,464,485,515), so the
kernel file is patched ON THE FLY in a temporary directory (``input.type()`` -> ``input.scalar_type()``) -- a
mechanical API rename that does not touch the arithmetic.  The resulting ``shift_cuda_ref*.so`` travels to the
GPU box (oracle/_ref/ is git-ignored, not gpurun-ignored) where tests/test_gpu_reference_ext.py uses it to pin
the oracle restatement and the product kernels against the reference's real kernels.
"""
import glob
import os
import shutil
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_CUDA_DIR = "/root/reference/model/Temporal_shift/cuda"
NAME = "shift_cuda_ref"


def built_library():
    hits = glob.glob(os.path.join(OUT, NAME + "*.so"))
    return hits[0] if hits else None


def build(force=False):
    if built_library() and not force:
        return built_library()
    if not os.path.isdir(REF_CUDA_DIR):
        raise FileNotFoundError(REF_CUDA_DIR)
    from torch.utils import cpp_extension
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="shift_ref_src_")
    try:
        cpp = os.path.join(tmp, "shift_cuda.cpp")
        cu = os.path.join(tmp, "shift_cuda_kernel.cu")
        shutil.copy(os.path.join(REF_CUDA_DIR, "shift_cuda.cpp"), cpp)
        with open(os.path.join(REF_CUDA_DIR, "shift_cuda_kernel.cu")) as f:
            src = f.read()
        with open(cu, "w") as f:
            f.write(src.replace("input.type()", "input.scalar_type()"))
        os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
        cpp_extension.load(name=NAME, sources=[cpp, cu], build_directory=OUT, verbose=False,
                           extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"], is_python_module=False)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    for junk in glob.glob(os.path.join(OUT, "*.o")) + glob.glob(os.path.join(OUT, "build.ninja")) + \
            glob.glob(os.path.join(OUT, ".ninja*")):
        os.remove(junk)
    return built_library()


def load():
    """import the compiled reference extension (GPU box: prebuilt file only)"""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    path = built_library()
    if path is None:
        raise FileNotFoundError("oracle/_ref/shift_cuda_ref*.so has not been built")
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build())
