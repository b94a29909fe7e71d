"""Builds libidrk.so (all CUDA kernels + the C ABI) for sm_100a, in-tree, with nvcc.

    python hashmodnffbanks-idr_b200/csrc/build.py [--force] [--verbose]

The .so stays next to the sources (git-ignored, but it travels to the GPU box with gpurun).
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libidrk.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"] + os.environ.get("IDRK_NVCC_FLAGS", "").split()


def sources():
    return sorted(glob.glob(os.path.join(HERE, "*.cu")))


def headers():
    return sorted(glob.glob(os.path.join(HERE, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "..", "include", "*.h")))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
    if not _stale(obj, [src] + headers() + [os.path.abspath(__file__)]):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    return obj, p.stderr if verbose else ""


def build(force=False, verbose=False):
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    srcs = sources()
    if force:
        for f in glob.glob(os.path.join(HERE, "build", "*.o")):
            os.remove(f)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in results]
    for _, log in results:
        if log:
            sys.stderr.write(log)
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
