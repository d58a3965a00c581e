#!/usr/bin/env python3
"""Every shipped ladder instance on device-resident random words, repeated: all runs must be bit-identical (no oracle:
repeatability only).  usage: ladder_stress.py [runs] [log2 lanes] [layout:instance filter]"""
import os, sys, json, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import ecsimd_b200
from ecsimd_b200 import device as dev, capi
R=int(sys.argv[1]) if len(sys.argv)>1 else 24
LOG2=int(sys.argv[2]) if len(sys.argv)>2 else 19
only=sys.argv[3] if len(sys.argv)>3 else ''
print('LIB', capi.LIB_PATH, 'R',R,'log2n',LOG2, flush=True)
ecsimd_b200.init(0)
n=1<<LOG2
g=torch.Generator(device='cuda'); g.manual_seed(1234)
def rnd(shape): return torch.randint(-2**31, 2**31-1, shape, dtype=torch.int32, device='cuda', generator=g)
def lanes_of(diff, layout):
    if layout=='lane': return torch.nonzero(diff.any(dim=1)).flatten()
    if layout=='pack4': return torch.nonzero(diff.any(dim=1)).flatten()*4
    return torch.nonzero(diff.any(dim=0).any(dim=1)).flatten()
total=0
for layout in ('pack4','lane','soa'):
    k=rnd(dev.empty(n,1,layout).shape); P=rnd(dev.empty(n,3,layout).shape)
    runs={'var':lambda o: dev.scalar_mult(o,k,P,n,layout), 'table':lambda o: dev.scalar_mult_base(o,k,n,layout,table=True), 'plain':lambda o: dev.scalar_mult_base(o,k,n,layout,table=False)}
    if layout=='soa':
        runs['var_nq']=lambda o: dev.scalar_mult(o,k,P,n,layout,quirk=False)
        runs['table_nq']=lambda o: dev.scalar_mult_base(o,k,n,layout,quirk=False,table=True)
        runs['plain_nq']=lambda o: dev.scalar_mult_base(o,k,n,layout,quirk=False,table=False)
    for name,run in runs.items():
        if only and only not in (layout+':'+name): continue
        ref=dev.empty(n,3,layout); run(ref); torch.cuda.synchronize()
        ev=[]
        for r in range(R):
            o=dev.empty(n,3,layout); o.zero_(); run(o); torch.cuda.synchronize()
            if not torch.equal(o,ref):
                l=lanes_of(o!=ref,layout).tolist()
                ev.append({'run':r,'nlanes':len(l),'lanes':l[:8]})
        total+=len(ev)
        print(json.dumps({'layout':layout,'inst':name,'runs':R,'lanes':n,'events':ev}),flush=True)
print('total events',total)
