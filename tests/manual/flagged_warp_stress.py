#!/usr/bin/env python3
"""Repeats the differential-fuzz inputs whose lanes take the out-of-line exact re-run (see README.md here and
profiles/r2e_recolor_bug/README.md).  usage: flagged_warp_stress.py [reps]"""
import os, sys, json, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import _libs, ecsimd_b200
from ecsimd_b200 import host as eng, capi, device as dev
REPS=int(sys.argv[1]) if len(sys.argv)>1 else 40
print('LIB', capi.LIB_PATH, flush=True)
ecsimd_b200.init(0)
orc=_libs.oracle(nt=os.cpu_count() or 1)
n=155648
G=np.concatenate([_libs.to_words([_libs.GX_INT]),_libs.to_words([_libs.GY_INT])],axis=1)
GJ=orc.from_affine(np.repeat(G,n,axis=0))
for s in (1,4,7):
    rnd=np.random.RandomState(1000+s)
    k=_libs.raw256(0xF00D0000+s,n)
    P=orc.from_affine(orc.to_affine(orc.scalar_mult(_libs.raw256(0xBEEF0000+s,n),GJ)))
    m=n//16
    P[:m,:16]=_libs.raw256(0xABCD0000+s,2*m).reshape(m,16)
    P[m:2*m,rnd.randint(0,16,size=m)]=0xFFFFFFFF
    k[:64]&=rnd.randint(0,2,size=(64,8)).astype(np.uint32)*np.uint32(0xFFFFFFFF)
    want=orc.scalar_mult(k,P)
    kp=eng.lane_to_pack4(k,1); Pp=eng.lane_to_pack4(P,3)
    assert np.array_equal(eng.pack4_to_lane(kp,1),k) and np.array_equal(eng.pack4_to_lane(Pp,3),P)
    wantp=eng.lane_to_pack4(want,3)
    ev=[]
    for rep in range(REPS):
        got=eng.scalar_mult(kp,Pp,layout='pack4')
        if not np.array_equal(got,wantp):
            bad=np.nonzero((eng.pack4_to_lane(got,3)!=want).any(axis=1))[0]
            ev.append({'rep':rep,'n':len(bad),'lanes':bad[:40].tolist()})
            np.save(os.path.join(ROOT,'gpurun_out','bad_got_s%d_r%d.npy'%(s,rep)), eng.pack4_to_lane(got,3)[bad])
            np.save(os.path.join(ROOT,'gpurun_out','bad_want_s%d_r%d.npy'%(s,rep)), want[bad])
    print(json.dumps({'seed':s,'path':'host pack4','reps':REPS,'events':ev}),flush=True)
    # device-resident, same data
    dk=torch.from_numpy(kp.view(np.int32)).cuda(); dP=torch.from_numpy(Pp.view(np.int32)).cuda(); dw=torch.from_numpy(wantp.view(np.int32)).cuda()
    ev=[]
    for rep in range(REPS):
        o=dev.empty(n,3,'pack4'); o.zero_(); dev.scalar_mult(o,dk,dP,n,'pack4'); torch.cuda.synchronize()
        if not torch.equal(o.reshape(-1),dw.reshape(-1)):
            bad=torch.nonzero((o.reshape(n//4,-1)!=dw.reshape(n//4,-1)).any(dim=1)).flatten()
            ev.append({'rep':rep,'npacks':len(bad),'packs':bad[:10].tolist()})
    print(json.dumps({'seed':s,'path':'device pack4','reps':REPS,'events':ev}),flush=True)
    # the other host calls of the fuzz
    ev=0
    wa=orc.to_affine(want)
    for rep in range(REPS//4):
        ev+=int((eng.scalar_mult_affine(k,P)!=wa).any(axis=1).sum())
    print(json.dumps({'seed':s,'path':'host lane affine','reps':REPS//4,'bad_lanes':ev}),flush=True)
