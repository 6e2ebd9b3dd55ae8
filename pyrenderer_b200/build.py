"""Build libprt.so (the sm_100a CUDA library behind include/prt.h) in-tree.

    python -m pyrenderer_b200.build [--force] [--verbose]

Plain nvcc, one object per .cu compiled in parallel, linked into
pyrenderer_b200/libprt.so.  No fast-math: IEEE division / sqrt are part of the
parity contract (csrc/intersect.cuh).
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libprt.so")
SOURCES = ["prt_api.cu", "traverse.cu", "bvh_build.cu", "wavefront.cu", "collective.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libprt.so cannot be built")
    return exe


def _newest_input():
    t = os.path.getmtime(os.path.join(HERE, "..", "include", "prt.h"))
    for f in os.listdir(CSRC):
        t = max(t, os.path.getmtime(os.path.join(CSRC, f)))
    return t


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: tuning variants (profiles/sweep.py), e.g. defines=("-DPRT_PSTACK=8",)."""
    global OBJ
    lib = out or LIB
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= _newest_input():
        return lib
    obj_dir = OBJ if out is None else OBJ + "_" + os.path.basename(out)
    os.makedirs(obj_dir, exist_ok=True)
    exe = nvcc()
    extra = (["-Xptxas", "-v"] if verbose else []) + list(defines)

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [exe] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        for src, obj, r in ex.map(compile_one, SOURCES):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"--- {src}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
            objs.append(obj)
    cmd = [exe, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libprt.so failed")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
