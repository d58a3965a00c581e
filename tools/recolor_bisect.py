#!/usr/bin/env python3
"""Bisection of a register renaming on the GPU (how the first miscompile of sass_recolor.py was located,
profiles/r2e_recolor_bug/): the difference between the cubin ptxas wrote and the re-coloured one is a set of register
changes; a change may need another one to vacate its target register, so the changes are grouped into strongly
connected components and ordered so that every prefix is a valid allocation.  `gen` writes the kernel's code for a
list of prefixes, `run` (on the GPU box) patches them into copies of the original cubin and runs each several times
with tools/variant_bench (checksum per variant + number of runs that differ from the first: an unstable variant
contains a race).

  recolor_bisect.py gen <kernel substr> <tag> auto:40 | lo:hi:n | k ...   -> build/bisect/<tag>_<k>.bin, <tag>.json
  recolor_bisect.py run <tag> [log2 lanes]                                -> one JSON line per variant
(copy build/obj/kernels_point.cu.keep/kernels_point.cubin.orig to build/bisect/ first: build/obj does not travel)
"""
import collections
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)


def gen(argv):
    sys.setrecursionlimit(100000)
    sys.path.insert(0, os.path.join("ecsimd_b200", "csrc"))
    import sass_recolor as rc
    ORIG = "build/obj/kernels_point.cu.keep/kernels_point.cubin.orig"
    NEW = os.environ.get("NEWCUBIN", "build/obj/kernels_point.cu.keep/kernels_point.cubin")
    SUB, TAG = argv[0], argv[1]
    os.makedirs("build/bisect", exist_ok=True)

    blob=open(ORIG,'rb').read(); nblob=open(NEW,'rb').read()
    secs=[n for n in rc.elf_sections(blob) if n.startswith('.text.') and SUB in n]; assert len(secs)==1
    sec=secs[0]
    _,off,ins=rc.disassemble(ORIG,blob,sec[6:],exact=True)
    A=rc.analyse(ins,False)
    _,off2,ins2=rc.disassemble(NEW,nblob,sec[6:],exact=True)
    for i in ins2: i.fields=rc.operand_fields(i)
    col0=[w['reg'] for w in A.webs]; col1=list(col0)
    for (k,oi,j),w in A.web_at.items(): col1[w]=ins2[k].fields[oi][0]+j
    D=set(w for w in range(len(A.webs)) if col0[w]!=col1[w])
    G=sorted(set(A.group_of[w] for w in D))
    succ={g:set() for g in G}
    for w in D:
        for o in A.adj[w]:
            if col1[w]==col0[o]:
                assert o in D
                if A.group_of[o]!=A.group_of[w]: succ[A.group_of[w]].add(A.group_of[o])
    # Tarjan
    index={}; low={}; st=[]; on=set(); sccs=[]; cnt=[0]
    def sc(v):
        index[v]=low[v]=cnt[0]; cnt[0]+=1; st.append(v); on.add(v)
        for x in succ[v]:
            if x not in index: sc(x); low[v]=min(low[v],low[x])
            elif x in on: low[v]=min(low[v],index[x])
        if low[v]==index[v]:
            c=[]
            while True:
                x=st.pop(); on.discard(x); c.append(x)
                if x==v: break
            sccs.append(c)
    for g in G:
        if g not in index: sc(g)
    print('groups changed',len(G),'sccs',len(sccs),'largest',max(len(c) for c in sccs))
    cuts=[]
    for a in argv[2:]:
        if a.startswith('auto:'):
            n=int(a[5:]); cuts+= [round(len(sccs)*i/n) for i in range(n+1)]
        elif ':' in a:
            lo,hi,n=map(int,a.split(':')); cuts+=[lo+round((hi-lo)*i/n) for i in range(n+1)]
        else: cuts.append(int(a))
    cuts=sorted(set(cuts))
    names=[]
    for m in cuts:
        col=list(col0)
        for c in sccs[:m]:
            for g in c:
                for w in A.groups[g]: col[w]=col1[w]
        for a in range(len(A.webs)):
            for b in A.adj[a]: assert col[a]!=col[b], ('improper', m)
        new,changed=rc.apply(blob,off,ins,A,col)
        fn='build/bisect/%s_%05d.bin'%(TAG,m)
        open(fn,'wb').write(new[off:off+len(ins)*16]); names.append(fn)
        print(m,changed)
    json.dump({'sec':sec,'off':off,'size':len(ins)*16,'files':names,'nsccs':len(sccs)},open('build/bisect/%s.json'%TAG,'w'))


def run(argv):
    tag = argv[0]
    log2n = argv[1] if len(argv) > 1 else "17"
    ORIG = "build/bisect/kernels_point.cubin.orig"
    meta = json.load(open("build/bisect/%s.json" % tag))
    blob = bytearray(open(ORIG, "rb").read())
    paths = []
    os.makedirs("/tmp/bis", exist_ok=True)
    for f in meta["files"]:
        b = bytearray(blob)
        code = open(f, "rb").read()
        assert len(code) == meta["size"]
        b[meta["off"]:meta["off"] + meta["size"]] = code
        p = "/tmp/bis/" + os.path.basename(f).replace(".bin", ".cubin")
        open(p, "wb").write(bytes(b))
        paths.append(p)
    env = dict(os.environ, KERNEL=meta["sec"][len(".text."):], REPEAT="6")
    if "ELi16E" in meta["sec"]:
        env["TABLE_KERNEL"] = "_ZN6ecb20018k_build_base_tableILb1EEEvP5uint4i"
    subprocess.run(["build/variant_bench", log2n] + paths, env=env)


if __name__ == "__main__":
    {"gen": gen, "run": run}[sys.argv[1]](sys.argv[2:])
