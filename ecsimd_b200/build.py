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


# Kernels whose registers are re-coloured after ptxas (csrc/sass_recolor.py: IMAD.WIDE multiplicands and ALU sources
# moved to different register banks; no instruction is added, removed or moved).  ECB200_RECOLOR=0 builds without the
# pass, ECB200_RECOLOR=strict makes a failure of the pass fatal (default: ship the kernels as ptxas wrote them and
# say so in recolor_report.json).  The default build only REPLAYS csrc/recolor_plans.json (the patches found and verified for
# the committed sources, keyed by the hash of ptxas' output: milliseconds, bit-reproducible); a kernel whose code has no
# stored patch (another ptxas) ships as ptxas wrote it and the report says so.  ECB200_RECOLOR=auto searches the missing
# ones (minutes per kernel), ECB200_RECOLOR=search ignores the stored patches and searches everything again.
RECOLOR = {"kernels_point.cu": ("k_scalar_mult_sync", "k_pointI", "k_to_affine", "k_from_x")}
RECOLOR_PLAN = os.path.join(CSRC, "recolor_plans.json")
RECOLOR_REPORT = os.path.join(HERE, "recolor_report.json")


def _compile_recolored(nvcc, src, obj, substr, verbose):
    """nvcc -c with one extra step between ptxas and fatbinary: replay nvcc's own sub-commands (nvcc -dryrun) and
    patch the .cubin in place."""
    import json
    import re
    keep = os.path.join(OBJDIR, os.path.basename(src) + ".keep")
    shutil.rmtree(keep, ignore_errors=True)
    os.makedirs(keep)
    dry = subprocess.run([nvcc] + NVCC_FLAGS + ["-dryrun", "--keep", "--keep-dir", keep, "-c", src, "-o", obj],
                         capture_output=True, text=True, check=True).stderr
    env = dict(os.environ)
    cmds = []
    for line in dry.splitlines():
        if not line.startswith("#$ "):
            continue
        line = line[3:]
        m = re.match(r"^([A-Za-z_][A-Za-z0-9_]*)=(.*)$", line)
        if m and not cmds:
            v = m.group(2).strip()
            env[m.group(1)] = os.path.expandvars(v.replace('"', "")) if m.group(1) != "PATH" else v
            continue
        cmds.append(line)
    env["PATH"] = env.get("PATH", os.environ["PATH"])
    report = None
    for c in cmds:
        if verbose:
            print(c[:160], flush=True)
        subprocess.run(["bash", "-c", c], check=True, env=env, cwd=ROOT)
        m = re.match(r'^ptxas .* -o "([^"]+\.cubin)"', c)
        if m and os.environ.get("ECB200_RECOLOR", "1") != "0":
            cubin = m.group(1)
            sys.path.insert(0, CSRC)
            try:
                import sass_recolor
                shutil.copyfile(cubin, cubin + ".orig")          # what ptxas wrote (tools/recolor_autotune.py starts from it)
                tmp = cubin + ".recolored"
                mode = os.environ.get("ECB200_RECOLOR", "1")
                report = sass_recolor.recolour_cubin(cubin, tmp, substr, plan_path=RECOLOR_PLAN, verbose=verbose,
                                                     use_plans=mode != "search", search_missing=mode in ("search", "auto", "strict"))
                os.replace(tmp, cubin)
            except Exception as e:
                if os.environ.get("ECB200_RECOLOR") == "strict":
                    raise
                print("WARNING: sass_recolor failed, shipping %s as ptxas wrote it: %s" % (substr, str(e)[:2000]), file=sys.stderr, flush=True)
                report = {"error": str(e)[:2000]}
            finally:
                sys.path.remove(CSRC)
    with open(RECOLOR_REPORT, "w") as f:
        json.dump({"source": os.path.basename(src), "kernels": report}, f, indent=1)


def build(force=False, verbose=False):
    regenerate()
    nvcc = _nvcc()
    os.makedirs(OBJDIR, exist_ok=True)
    deps = [os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        if force or _newer(o, [s] + deps + ([os.path.join(CSRC, "sass_recolor.py")] if src in RECOLOR else [])):
            if src in RECOLOR and os.environ.get("ECB200_RECOLOR", "1") != "0":
                jobs.append(("recolor", s, o, RECOLOR[src]))
            else:
                jobs.append([nvcc] + NVCC_FLAGS + ["-c", s, "-o", o])
    def run(cmd):
        if isinstance(cmd, tuple):
            return _compile_recolored(nvcc, cmd[1], cmd[2], cmd[3], verbose)
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
