"""Offline build of the sm_100a C-ABI library (replaces the reference's setup.py,
model/Temporal_shift/cuda/setup.py:4-14, which builds for whatever arch is visible).

``build()`` compiles every ``csrc/*.cu`` with nvcc for ``sm_100a`` only (no PTX fallback, no other arch) and
links ``shiftgcn_b200/lib/libshiftgcn_b200.so`` in-tree, so the binary travels with the repository snapshot.
"""
import concurrent.futures
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libshiftgcn_b200.so")
OBJ_DIR = os.path.join(LIB_DIR, "obj")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(os.path.dirname(_HERE), "include"),
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; shiftgcn_b200 needs the CUDA 12.9 toolkit to build")
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(os.path.dirname(_HERE), "include", "shiftgcn_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, verbose):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), _headers_mtime()):
        return obj, False
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return obj, True


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    srcs = sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as pool:
        results = list(pool.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in results]
    rebuilt = any(r for _, r in results)
    if rebuilt or not os.path.exists(LIB_PATH):
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs, "-lcudart"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(verbose=True))
