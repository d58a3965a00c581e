"""CPU: the multi-GPU host logic (index-range sharding, seeded generation consistency,
max-over-ranks reduction) with world_size = 2 over gloo.  The compute on each "rank" is
the CPU oracle standing in for the GPU kernel: what is tested is that the union of the
shards reproduces the single-rank batch bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition():
    from ecsimd_b200.shard import shard_range
    for n in (0, 1, 127, 128, 129, 1000, 1 << 20, (1 << 26) + 5):
        for world in (1, 2, 3, 4, 8):
            pieces = [shard_range(n, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            for (a, b), (c, d) in zip(pieces, pieces[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in pieces]
            assert all(lo % 128 == 0 for lo, _ in pieces if lo < n)
            assert max(sizes) - min(sizes) < 256


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import _libs
    from ecsimd_b200 import shard
    r, w, _ = shard.init_distributed("gloo")
    lo, hi = shard.shard_range(n, r, w)
    orc = _libs.oracle(1)
    a = _libs.field_elems(0xEC51D001, hi - lo, start=lo)
    b = _libs.field_elems(0xEC51D002, hi - lo, start=lo)
    out = orc.mgry_mul(a, b)
    chk = np.bitwise_xor.reduce(out.reshape(-1, 8), axis=0).astype(np.uint32)
    shard.barrier()
    total = shard.xor_over_ranks(chk)
    tmax = shard.max_over_ranks(1.0 + r)
    lanes = shard.sum_over_ranks(hi - lo)
    q.put((r, lo, hi, total.tolist(), tmax, lanes))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_single_rank():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _libs
    n, world = 1000, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    orc = _libs.oracle(1)
    full = orc.mgry_mul(_libs.field_elems(0xEC51D001, n), _libs.field_elems(0xEC51D002, n))
    want = np.bitwise_xor.reduce(full, axis=0).astype(np.uint32).tolist()
    res.sort()
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n
    for r in res:
        assert r[3] == want          # union of shards == whole batch
        assert r[4] == 2.0           # max over ranks of (1 + rank)
        assert r[5] == n
