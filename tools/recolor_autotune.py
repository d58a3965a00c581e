#!/usr/bin/env python3
"""Several re-colourings of ONE shipped ladder kernel (different seeds of sass_recolor's search), each written as a
complete kernels_point.cubin variant, to be timed on a GPU with tools/variant_bench.cu (KERNEL=<name>); `pick` then
stores the fastest one's patch in csrc/recolor_plans.json.  The cost model of the search is fitted to +-0.2 %, so
the variants differ by a few tenths of a percent on the device: this closes that gap by measurement.

  recolor_autotune.py gen  <kernel substring> <n seeds>      -> build/autotune/<tag>_s<seed>.cubin (+ list.txt)
  recolor_autotune.py pick <kernel substring> <results.jsonl> -> updates recolor_plans.json (rebuild replays it)
"""
import json
import multiprocessing
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ecsimd_b200", "csrc")
sys.path.insert(0, CSRC)
import sass_recolor as rc  # noqa: E402

CUBIN = os.path.join(ROOT, "build", "obj", "kernels_point.cu.keep", "kernels_point.cubin.orig")
OUT = os.path.join(ROOT, "build", "autotune")
PLANS = os.path.join(CSRC, "recolor_plans.json")


def _one(args):
    sec, seed, tag = args
    r = rc.recolour_section(CUBIN, sec, iters=30000, seed=seed)
    blob = bytearray(open(CUBIN, "rb").read())
    blob[r["offset"]:r["offset"] + len(r["code"])] = r["code"]
    path = os.path.join(OUT, "%s_s%d.cubin" % (tag, seed))
    open(path, "wb").write(bytes(blob))
    return path, r["cost_after"]


def main():
    mode, sub = sys.argv[1], sys.argv[2]
    blob = open(CUBIN, "rb").read()
    secs = [n for n in rc.elf_sections(blob) if n.startswith(".text.") and sub in n]
    assert len(secs) == 1, secs
    sec = secs[0]
    tag = "".join(ch for ch in sub if ch.isalnum())[-24:]
    os.makedirs(OUT, exist_ok=True)
    if mode == "gen":
        n = int(sys.argv[3])
        with multiprocessing.Pool(min(n, os.cpu_count() or 1)) as pool:
            res = pool.map(_one, [(sec, 100 + s, tag) for s in range(n)], chunksize=1)
        for p, c in res:
            print(p, round(c, 2))
        open(os.path.join(OUT, "%s_list.txt" % tag), "w").write(" ".join(os.path.relpath(p, ROOT) for p, _ in res))
        print("KERNEL=%s" % sec[len(".text."):])
    else:
        best = None
        for l in open(sys.argv[3]):
            if '"variant"' in l:
                d = json.loads(l)
                if tag in d["variant"] and (best is None or d["ms"] < best["ms"]):
                    best = d
        print("best:", best)
        off, size, _ = rc.elf_sections(blob)[sec]
        code = blob[off:off + size]
        new = open(os.path.join(ROOT, best["variant"]), "rb").read()[off:off + size]
        plans = json.load(open(PLANS))
        plans[sec].update({"key": rc.code_hash(code), "patched_key": rc.code_hash(new), "xor": rc._pack(bytes(a ^ b for a, b in zip(code, new))),
                           "autotuned_ms_2p20": best["ms"]})
        json.dump(plans, open(PLANS, "w"), indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
