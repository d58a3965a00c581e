"""GPU: BASELINE configs[1] (2^22 points, add + double) and configs[3] (2^24 scalars on the
generator) at full size, device resident: oracle on sampled lanes plus size-independent properties
(co-Z invariants, two routes to the same point, determinism of the order-independent checksum)."""
import numpy as np
import pytest

import _libs
from _libs import GX_INT, GY_INT, raw256, to_words

pytestmark = pytest.mark.gpu


def _lanes(t, idx, nc):
    """gather lanes `idx` of a planar (2*nc, n, 4) device tensor into a (len(idx), 8*nc) numpy array"""
    import torch
    sel = t[:, torch.as_tensor(idx, device=t.device), :].cpu().numpy().view(np.uint32)     # (2nc, m, 4)
    m = len(idx)
    return np.ascontiguousarray(sel.reshape(nc, 2, m, 4).transpose(2, 0, 1, 3)).reshape(m, 8 * nc)


def test_point_add_double_2p22(eng, orc):
    import torch
    from ecsimd_b200 import device as dev
    n = 1 << 22
    r = dev.synth_values(dev.empty(n, 1), 0xEC51D003, 0, n, 0)
    P = dev.from_affine(dev.empty(n, 3), dev.to_affine(dev.empty(n, 2), dev.scalar_mult_base(dev.empty(n, 3), r, n), n), n)
    P1, T3 = dev.empty(n, 3), dev.empty(n, 3)
    dev.trplu(P1, T3, P, n)                               # (P rescaled, 3P)
    Q, F = dev.empty(n, 3), dev.empty(n, 3)
    dev.zdau(Q, F, T3, P1, n)                             # 2*(3P) + P = 7P, P rescaled again
    torch.cuda.synchronize()
    # co-Z invariants on every lane: the rewritten operand and the result share Z
    assert torch.equal(P1[4:6], T3[4:6]) and torch.equal(Q[4:6], F[4:6])
    idx = np.concatenate([np.arange(0, 256), np.arange(n - 256, n), np.arange(0, n, 65521)])
    Ps = _lanes(P, idx, 3)
    w1, w3 = orc.trplu(Ps)
    assert np.array_equal(_lanes(P1, idx, 3), w1) and np.array_equal(_lanes(T3, idx, 3), w3)
    wq, wf = orc.zdau(w3, w1)
    assert np.array_equal(_lanes(Q, idx, 3), wq) and np.array_equal(_lanes(F, idx, 3), wf)
    # 7P by another route (scalar_mult with k = 7) has the same affine coordinates on every lane
    k7 = torch.zeros((2, n, 4), dtype=torch.int32, device=P.device); k7[0, :, 0] = 7
    S = dev.scalar_mult(dev.empty(n, 3), k7, P, n)
    a1 = dev.to_affine(dev.empty(n, 2), F, n)
    a2 = dev.to_affine(dev.empty(n, 2), S, n)
    torch.cuda.synchronize()
    bad = (a1 != a2).any(dim=0).any(dim=1).sum().item()
    assert bad <= 64, bad          # only lanes hit by the reference's squaring defect may differ


def test_generator_2p24(eng, orc):
    import torch
    from ecsimd_b200 import device as dev
    n = 1 << 24
    k = dev.synth_values(dev.empty(n, 1), 0xEC51D004, 0, n, 0)
    out = dev.scalar_mult_base(dev.empty(n, 3), k, n)
    torch.cuda.synchronize()
    idx = np.concatenate([np.arange(0, 128), np.arange(n - 128, n), np.arange(0, n, 262139)])
    ks = _lanes(k, idx, 1)
    assert np.array_equal(ks, np.concatenate([raw256(0xEC51D004, 1, start=int(i)) for i in idx]))   # seeded stream
    G = np.concatenate([to_words([GX_INT]), to_words([GY_INT])], axis=1)
    GJ = orc.from_affine(np.repeat(G, len(idx), axis=0))
    assert np.array_equal(_lanes(out, idx, 3), orc.scalar_mult(ks, GJ))
    # determinism and order independence: second run, and the two halves' checksums, fold to the same words
    c1 = dev.checksum(out)
    out2 = dev.scalar_mult_base(dev.empty(n, 3), k, n)
    assert np.array_equal(dev.checksum(out2), c1)
    assert torch.equal(out, out2)
    # the fixed-base table (all 2^16 indices occur among 2^24 random scalars) against the plain ladder
    out3 = dev.scalar_mult_base(dev.empty(n, 3), k, n, table=False)
    assert torch.equal(out, out3)


def test_config5_shards_fold_to_the_whole(eng, orc):
    """BASELINE configs[4] on one GPU: the index range of a 2^23-lane variable-base batch (the share
    of one GPU in the 2^26-lane, 8-GPU job) cut by shard.shard_range into 8 'ranks' that regenerate
    their inputs from (seed, global index): every shard's output equals the same lanes of the
    whole-batch run, the xor of the shard checksums is the checksum of the whole, and sampled lanes
    match the oracle."""
    import torch
    from ecsimd_b200 import device as dev
    from ecsimd_b200.shard import shard_range
    n, world = 1 << 23, 8

    def inputs(lo, m):
        k = dev.synth_values(dev.empty(m, 1), 0xEC51D004, lo, m, 0)
        r = dev.synth_values(dev.empty(m, 1), 0xEC51D003, lo, m, 0)
        J = dev.scalar_mult_base(dev.empty(m, 3), r, m)
        P = dev.from_affine(dev.empty(m, 3), dev.to_affine(dev.empty(m, 2), J, m), m)
        return k, P

    k, P = inputs(0, n)
    whole = dev.scalar_mult(dev.empty(n, 3), k, P, n)
    torch.cuda.synchronize()
    total = np.zeros(8, np.uint32)
    for rank in range(world):
        lo, hi = shard_range(n, rank, world)
        ks, Ps = inputs(lo, hi - lo)
        assert torch.equal(ks, k[:, lo:hi]) and torch.equal(Ps, P[:, lo:hi])          # seeded regeneration
        part = dev.scalar_mult(dev.empty(hi - lo, 3), ks, Ps, hi - lo)
        torch.cuda.synchronize()
        assert torch.equal(part, whole[:, lo:hi])
        total ^= dev.checksum(part)
    assert np.array_equal(total, dev.checksum(whole))
    idx = np.concatenate([np.arange(0, 64), np.arange(n - 64, n), np.arange(0, n, 1048573)])
    assert np.array_equal(_lanes(whole, idx, 3), orc.scalar_mult(_lanes(k, idx, 1), _lanes(P, idx, 3)))
