"""Build libmms_b200.so (hand-written CUDA for sm_100a + the C-ABI) in-tree with nvcc.

    python -m mms_answer_selection_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box
with the repo snapshot; nothing is JIT-compiled at run time.
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libmms_b200.so")

NVCC = os.environ.get("MMS_NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = os.environ.get("MMS_HOST_CXX", "/usr/bin/g++")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "-Xcudafe", "--diag_suppress=177",
] + os.environ.get("MMS_NVCC_EXTRA", "").split()


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "tc", "*.cu")))


def headers():
    return (glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "tc", "*.cuh")) +
            glob.glob(os.path.join(HERE, "..", "include", "*.h")))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, obj, verbose):
    cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = headers() + [os.path.abspath(__file__)]
    jobs = []
    objs = []
    for src in sources():
        rel = os.path.relpath(src, CSRC).replace(os.sep, "_")
        obj = os.path.join(OBJ, rel[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append((src, obj))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, rc, out in ex.map(lambda j: _compile(j[0], j[1], verbose), jobs):
                if verbose or rc != 0:
                    sys.stderr.write(out)
                if rc != 0:
                    raise RuntimeError("nvcc failed on %s" % src)
    if jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-ccbin", HOST_CXX, "-o", LIB] + objs + ["-lcuda"]
        # libcuda is only needed for cuTensorMapEncodeTiled, resolved through
        # cudaGetDriverEntryPoint at run time -> do not link it (no driver on build hosts)
        cmd = cmd[:-1]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib)
