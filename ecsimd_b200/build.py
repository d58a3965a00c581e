"""In-tree build of the CUDA engine: nvcc -> ecsimd_b200/libecb200.so (sm_100a only).

The shared library is the product: a C-ABI (include/ecb200.h) over hand-written
CUDA kernels.  It is built in-tree so that it travels to the GPU box with the
repository snapshot; it is git-ignored (*.so).
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OBJDIR = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(HERE, "libecb200.so")

SOURCES = ["kernels_field.cu", "kernels_point.cu", "kernels_generic.cu"]
HEADERS = ["fp256.cuh", "fp256_mul_gen.cuh", "fpgen.cuh", "point.cuh", "zdau_order.inc", "layout.cuh", "host_common.cuh", os.path.join("..", "..", "include", "ecb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-diag-suppress", "550"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA engine cannot be built (there is no CPU fallback)")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def regenerate():
    """Re-run the multiplier generator (also re-checks the schedule in simulation)."""
    gen = os.path.join(CSRC, "gen_fp256.py")
    out = os.path.join(CSRC, "fp256_mul_gen.cuh")
    if _newer(out, [gen]):
        subprocess.run([sys.executable, gen], check=True, cwd=CSRC)


def build(force=False, verbose=False):
    regenerate()
    nvcc = _nvcc()
    os.makedirs(OBJDIR, exist_ok=True)
    deps = [os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        if force or _newer(o, [s] + deps):
            jobs.append([nvcc] + NVCC_FLAGS + ["-c", s, "-o", o])
    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJDIR, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _newer(LIB, objs):
        run([nvcc, "-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC", "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
