#!/usr/bin/env python3
"""Audit of a built library's re-coloured kernels against the CURRENT model of sass_recolor.py, independent of the
search that produced them: for every re-coloured kernel of build/obj/kernels_point.cu.keep/kernels_point.cubin
(next to the .orig that ptxas wrote) it re-derives the in-use windows (scoreboard shadows, dead destinations, reuse
operands) on both versions, by register number, and lists every redefinition inside a window that the original does
not have at the same place; it also counts the register fields changed outside the hot loop.
usage: tools/recolor_audit.py [kernel substring ...]"""
import os, sys, json
from concurrent.futures import ProcessPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ecsimd_b200", "csrc"))
import sass_recolor as rc  # noqa: E402
KEEP = os.path.join(ROOT, "build", "obj", "kernels_point.cu.keep")
ORIG, NEW = os.path.join(KEEP, "kernels_point.cubin.orig"), os.path.join(KEEP, "kernels_point.cubin")


def audit(sec):
    def load(path):
        blob = open(path, "rb").read()
        _, off, ins = rc.disassemble(path, blob, sec[len(".text."):], exact=True)
        A = rc.analyse(ins)
        return ins, A
    a, Aa = load(ORIG)
    b, Ab = load(NEW)
    new = sorted(rc.hidden_hazards(b, Ab.lout) - rc.hidden_hazards(a, Aa.lout))
    rng = rc.hot_range(a)
    rc.mark_hot(a, rng, Aa.calls)
    changed_hot = changed_cold = 0
    for x, y in zip(a, b):
        for fx, fy in zip(x.fields, y.fields):
            if fx[0] != fy[0]:
                if x.hot:
                    changed_hot += 1
                else:
                    changed_cold += 1
    return {"kernel": sec[len(".text."):], "instructions": len(a), "new_hazards": ["%04x %s <- %04x %s" % (b[k].addr, b[k].text, b[j].addr, b[j].text) for k, oi, o, j in new[:8]],
            "n_new_hazards": len(new), "fields_changed_hot": changed_hot, "fields_changed_elsewhere": changed_cold}


def main():
    blob = open(ORIG, "rb").read()
    subs = sys.argv[1:] or ["k_scalar_mult_sync", "k_pointI", "k_to_affine", "k_from_x"]
    secs = sorted(n for n in rc.elf_sections(blob) if n.startswith(".text.") and any(s in n for s in subs))
    bad = 0
    with ProcessPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for r in ex.map(audit, secs):
            print(json.dumps(r), flush=True)
            bad += r["n_new_hazards"]
    print(json.dumps({"kernels": len(secs), "new_hazards_total": bad}))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
