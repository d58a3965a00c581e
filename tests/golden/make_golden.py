#!/usr/bin/env python3
"""Generate tests/golden/* from the UNMODIFIED reference compiled in this container
(oracle/_ref/libecsimd_ref.so, built by oracle/Makefile from /root/reference).

Run here (the reference does not exist on the GPU box); the outputs are committed:
  field_ops.npz           inputs a,b (random canonical + edge + squaring-quirk values) and the
                          reference's mgry_add/sub/mul/sqr/shl1/opposite/from/to_classical,
                          mul512, square512
  point_ops.npz           points P, scalars k and the reference's DBLU/ZADDU/ZDAU/ADD_Z2_1/
                          scalar_mult/to_affine outputs (Jacobian, Montgomery form)
  quirk_scalar_mult.json  (k, P) pairs whose ladder hits the reference's lost-carry squaring
                          defect (found by a search with the oracle's wrap counter, outputs
                          taken from the reference)
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _libs  # noqa: E402
from _libs import EDGE_FIELD, EDGE_SCALARS, GX_INT, GY_INT, QUIRK_FIELD, field_elems, raw256, to_ints, to_words  # noqa: E402


def main(search_lanes):
    ref = _libs.reference(nt=os.cpu_count())
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    orc = _libs.oracle(nt=os.cpu_count())

    edge = to_words(EDGE_FIELD + QUIRK_FIELD)
    a = np.concatenate([np.repeat(edge, len(edge), axis=0), field_elems(0xEC51D001, 256), _libs.quirk_stress(64, 5)])
    b = np.concatenate([np.tile(edge, (len(edge), 1)), field_elems(0xEC51D002, 256), field_elems(0xEC51D006, 64)])
    out = {"a": a, "b": b}
    for op in ("mgry_add", "mgry_sub", "mgry_mul"):
        out[op] = getattr(ref, op)(a, b)
    for op in ("mgry_sqr", "mgry_shl1", "opposite", "from_classical", "to_classical"):
        out[op] = getattr(ref, op)(a)
    out["mul512"] = ref.mul512(a, b)
    out["square512"] = ref.square512(a)
    np.savez_compressed(os.path.join(HERE, "field_ops.npz"), **out)

    n = 96
    G = np.concatenate([to_words([GX_INT]), to_words([GY_INT])], axis=1)
    GJ = ref.from_affine(np.repeat(G, n, axis=0))
    P = ref.from_affine(ref.to_affine(ref.scalar_mult(raw256(0xEC51D003, n), GJ)))
    P[0] = GJ[0]
    k = raw256(0xEC51D004, n)
    for i, v in enumerate(EDGE_SCALARS):
        k[i] = to_words([v])[0]
    pt = {"P": P, "k": k}
    pt["dblu_P"], pt["dblu_2P"] = ref.dblu(P)
    pt["zaddu_P"], pt["zaddu_R"] = ref.zaddu(pt["dblu_P"], pt["dblu_2P"])
    pt["zdau_Q"], pt["zdau_R"] = ref.zdau(pt["zaddu_R"], pt["zaddu_P"])
    pt["add_z2_1"] = ref.add_z2_1(pt["zdau_R"], P)
    pt["scalar_mult"] = ref.scalar_mult(k, P)
    pt["to_affine"] = ref.to_affine(pt["scalar_mult"])
    np.savez_compressed(os.path.join(HERE, "point_ops.npz"), **pt)

    # ---- search for ladder runs that hit the squaring defect --------------------------------
    found = {"k": [], "Px": [], "Py": [], "X": [], "Y": [], "Z": [], "wraps": []}
    f = orc.lib.orc_scalar_mult_wraps
    f.restype = None
    base = ref.from_affine(ref.to_affine(ref.scalar_mult(raw256(0xEC51D007, 1024), np.repeat(GJ[:1], 1024, axis=0))))
    chunk = 1 << 16
    for c in range(0, search_lanes, chunk):
        kk = raw256(0xEC51D008, chunk, start=c)
        PP = np.tile(base, (chunk // 1024, 1))
        o = np.zeros((chunk, 24), np.uint32)
        w = np.zeros(chunk, np.uint32)
        f(o.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), kk.ctypes.data_as(C.c_void_p),
          PP.ctypes.data_as(C.c_void_p), C.c_size_t(chunk), C.c_int(os.cpu_count()))
        for i in np.nonzero(w)[0]:
            r = ref.scalar_mult(kk[i:i + 1].repeat(4, axis=0), PP[i:i + 1].repeat(4, axis=0))[0]
            assert np.array_equal(r, o[i]), "oracle and reference disagree on a quirk lane"
            found["k"].append("%064x" % to_ints(kk[i])[0])
            found["Px"].append("%064x" % to_ints(PP[i, :8])[0])
            found["Py"].append("%064x" % to_ints(PP[i, 8:16])[0])
            for cname, sl in (("X", slice(0, 8)), ("Y", slice(8, 16)), ("Z", slice(16, 24))):
                found[cname].append("%064x" % to_ints(r[sl])[0])
            found["wraps"].append(int(w[i]))
        print("searched %d lanes, found %d" % (c + chunk, len(found["k"])), flush=True)
        if len(found["k"]) >= 4:
            break
    if found["k"]:
        json.dump(found, open(os.path.join(HERE, "quirk_scalar_mult.json"), "w"), indent=1)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21)
