"""GPU: the host-memory batch path of the scalar multiplication -- pinned and pageable caller buffers
(the library's bounce buffers), single-process multi-GPU dispatch (ecb200_init_devices: the reference's
caller is one process, benchs/curve_group.cpp:23-60), and the results of the CUDA path compared DIRECTLY
with the fixtures generated from the compiled reference (tests/golden/*.npz), without the C restatement
in between."""
import json
import os

import numpy as np
import pytest

import _libs
from _libs import GX_INT, GY_INT, raw256, to_words

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _points(eng, n, seed=0xEC51D003):
    """n Jacobian points r_i * G with Z = R (lane layout), computed by the engine itself"""
    return eng.from_affine(eng.to_affine(eng.scalar_mult_base(raw256(seed, n))))


# ---- straight against the reference's own outputs ---------------------------------------------------
def test_gpu_vs_reference_golden_field(eng):
    g = np.load(os.path.join(GOLDEN, "field_ops.npz"))
    a, b = g["a"], g["b"]
    assert np.array_equal(eng.mgry_add(a, b), g["mgry_add"])
    assert np.array_equal(eng.mgry_sub(a, b), g["mgry_sub"])
    assert np.array_equal(eng.mgry_mul(a, b), g["mgry_mul"])
    assert np.array_equal(eng.mgry_sqr(a), g["mgry_sqr"])
    assert np.array_equal(eng.mgry_shift_left(a, 1), g["mgry_shl1"])
    assert np.array_equal(eng.opposite(a), g["opposite"])
    assert np.array_equal(eng.from_classical(a), g["from_classical"])
    assert np.array_equal(eng.to_classical(a), g["to_classical"])
    assert np.array_equal(eng.mul512(a, b), g["mul512"])
    assert np.array_equal(eng.square512(a), g["square512"])


def test_gpu_vs_reference_golden_points(eng):
    g = np.load(os.path.join(GOLDEN, "point_ops.npz"))
    P, k = g["P"], g["k"]
    for layout in ("lane", "soa"):
        cv = (lambda x, nc: x) if layout == "lane" else (lambda x, nc: eng.lane_to_soa(x, nc))
        bk = (lambda x, nc: x) if layout == "lane" else (lambda x, nc: eng.soa_to_lane(x, nc))
        p1, d = eng.DBLU(cv(P, 3), layout=layout)
        assert np.array_equal(bk(p1, 3), g["dblu_P"]) and np.array_equal(bk(d, 3), g["dblu_2P"])
        p2, t = eng.ZADDU(p1, d, layout=layout)
        assert np.array_equal(bk(p2, 3), g["zaddu_P"]) and np.array_equal(bk(t, 3), g["zaddu_R"])
        q, r = eng.ZDAU(t, p2, layout=layout)
        assert np.array_equal(bk(q, 3), g["zdau_Q"]) and np.array_equal(bk(r, 3), g["zdau_R"])
        assert np.array_equal(bk(eng.ADD_Z2_1(r, cv(P, 3), layout=layout), 3), g["add_z2_1"])
        sm = eng.scalar_mult(cv(k, 1), cv(P, 3), layout=layout)
        assert np.array_equal(bk(sm, 3), g["scalar_mult"])
        assert np.array_equal(bk(eng.to_affine(sm, layout=layout), 2), g["to_affine"])
    assert np.array_equal(eng.scalar_mult_affine(k, P), g["to_affine"])


# ---- caller-side memory kinds ----------------------------------------------------------------------------
def test_affine_call_keeps_converted_point_array_alive(eng, orc):
    """ecsimd_b200.host.scalar_mult_affine with a P that needs conversion (int32 view, non-contiguous
    slice): the converted temporary must stay alive for the whole C call"""
    n = 64
    k = raw256(0xEC51D004, n)
    P = _points(eng, n)
    want = orc.to_affine(orc.scalar_mult(k, P))
    wide = np.zeros((n, 48), np.int32)
    wide[:, ::2] = P.view(np.int32)
    for _ in range(8):     # freed temporaries get recycled quickly: repeat
        junk = [np.full((n, 24), 0x5A5A5A5A, np.uint32) for _ in range(4)]
        assert np.array_equal(eng.scalar_mult_affine(k, wide[:, ::2]), want)
        del junk


def test_pageable_and_pinned_host_batches_agree(eng, orc):
    """a batch of several chunks (the pipelined path) from pageable numpy buffers (bounce buffers) and from
    pinned torch buffers, pack4 layout: identical to each other, sampled lanes identical to the oracle"""
    import torch
    from ecsimd_b200 import capi
    n = 3 * 148 * 512 + 4 * 37          # three full chunks and a ragged one
    k = raw256(0xEC51D004, n)
    P = _points(eng, n)
    kp, Pp = eng.lane_to_pack4(k, 1), eng.lane_to_pack4(P, 3)
    flags = capi.LAYOUT_PACK4 | capi.MEM_HOST
    out_pageable = np.zeros((n // 4, 96), np.uint32)
    capi.call("ecb200_scalar_mult_p256", capi._p(out_pageable), capi._p(kp), capi._p(Pp), n, flags, None)
    hk = torch.from_numpy(kp.view(np.int32)).pin_memory()
    hP = torch.from_numpy(Pp.view(np.int32)).pin_memory()
    hout = torch.zeros((n // 4, 96), dtype=torch.int32).pin_memory()
    capi.call("ecb200_scalar_mult_p256", hout.data_ptr(), hk.data_ptr(), hP.data_ptr(), n, flags, None)
    assert np.array_equal(out_pageable, hout.numpy().view(np.uint32))
    got = eng.pack4_to_lane(out_pageable, 3)
    idx = np.unique(np.concatenate([np.arange(64), np.arange(n - 64, n), np.arange(0, n, 4099)]))
    assert np.array_equal(got[idx], orc.scalar_mult(k[idx], P[idx]))
    # the fused affine call on the same pageable buffers
    xy = np.zeros((n // 4, 64), np.uint32)
    capi.call("ecb200_scalar_mult_p256_affine", capi._p(xy), capi._p(kp), capi._p(Pp), n, flags, None)
    assert np.array_equal(eng.pack4_to_lane(xy, 2)[idx], orc.to_affine(orc.scalar_mult(k[idx], P[idx])))


def test_single_process_multi_device_dispatch(eng, orc):
    """ecb200_init_devices over every visible device (1 on the single-GPU box: the dispatch then falls back to the
    one-device pipeline): the host batch is cut over the devices and folds back to the single-device result"""
    import torch
    import ecsimd_b200
    from ecsimd_b200 import capi
    ndev = torch.cuda.device_count()
    n = 2 * 148 * 512 * max(1, ndev) + 512 * 3 + 8
    k = raw256(0xEC51D004, n)
    P = _points(eng, n)
    flags = capi.LAYOUT_LANE | capi.MEM_HOST
    single = np.zeros((n, 24), np.uint32)
    capi.call("ecb200_scalar_mult_p256", capi._p(single), capi._p(k), capi._p(P), n, flags, None)
    try:
        ecsimd_b200.init_devices(list(range(ndev)))
        assert ecsimd_b200.device_count() == max(1, ndev)
        multi = np.zeros((n, 24), np.uint32)
        capi.call("ecb200_scalar_mult_p256", capi._p(multi), capi._p(k), capi._p(P), n, flags, None)
        base_multi = np.zeros((n, 24), np.uint32)
        capi.call("ecb200_scalar_mult_p256_base", capi._p(base_multi), capi._p(k), n, flags, None)
    finally:
        ecsimd_b200.init_devices([])
        ecsimd_b200.init(0)
    assert ecsimd_b200.device_count() == 1
    assert np.array_equal(multi, single)
    idx = np.unique(np.concatenate([np.arange(32), np.arange(n - 32, n), np.arange(0, n, 9973)]))
    assert np.array_equal(multi[idx], orc.scalar_mult(k[idx], P[idx]))
    G = np.concatenate([to_words([GX_INT]), to_words([GY_INT])], axis=1)
    assert np.array_equal(base_multi[idx], orc.scalar_mult(k[idx], orc.from_affine(np.repeat(G, len(idx), axis=0))))


def test_shutdown_releases_and_recovers(eng, orc):
    """ecb200_shutdown gives the pool, tables, streams and bounce buffers back; the next call re-creates them"""
    import ecsimd_b200
    n = 256
    k = raw256(0xEC51D004, n)
    a = eng.scalar_mult_base(k)
    ecsimd_b200.shutdown()
    b = eng.scalar_mult_base(k)        # lazily re-created context and table
    ecsimd_b200.init(0)
    assert np.array_equal(a, b)
