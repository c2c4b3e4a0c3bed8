"""Build libsb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m sparsebase_b200.build [--force]

Each .cu under csrc/ is compiled to an object in csrc/_build/ (in parallel) and linked into
sparsebase_b200/libsb200.so.  The .so is git-ignored but travels to the GPU box with gpurun.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsb200.so")
NVCC = os.environ.get("SB200_NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = os.environ.get("SB200_HOST_CXX", "/usr/bin/g++")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr", "-Xptxas", "-v" if os.environ.get("SB200_PTXAS_V") else "-O3",
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [
        os.path.join(HERE, "..", "include", "sb200.h")]
    bdir = os.path.join(CSRC, "_build")
    os.makedirs(bdir, exist_ok=True)
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(bdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [NVCC] + FLAGS + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for s, r in ex.map(compile_one, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(f"--- nvcc {os.path.basename(s)} ---\n{r.stdout}{r.stderr}\n")
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {s}")
    if jobs or force or _newer(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-ccbin", HOST_CXX, "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
