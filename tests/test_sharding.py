"""N>1 host logic on CPU: world_size-2 gloo run of the sharding helpers (no GPU).  The data path has no
collective; what is checked is that the shards tile the batch in rank order and that the oracle result of
shard r equals the slice of the single-rank result (SURVEY.md §4 "multi-GPU test")."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from homomorph_rust_b200.sharding import gather_plaintexts, max_over_ranks, shard, shard_range


def test_shard_range_tiles():
    for n in [0, 1, 7, 8, 9, 2**18, 2**18 + 5]:
        for world in [1, 2, 4, 8]:
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import hmoracle as orc

    d, dp, delta, tau, L = 64, 32, 1, 32, 8
    rng = np.random.default_rng(99)  # same stream on every rank: same keys, same full batch
    sk, pk = orc.keygen(d, dp, delta, tau, rng)
    n = 11
    a = rng.integers(0, 256, size=n, dtype=np.uint8)
    b = rng.integers(0, 256, size=n, dtype=np.uint8)
    ma = rng.integers(0, 256, size=n * L * 4, dtype=np.uint8)
    mb = rng.integers(0, 256, size=n * L * 4, dtype=np.uint8)

    def run(va, m1, vb, m2):
        ca = orc.encrypt(pk, va, 1, m1)[0]
        cb = orc.encrypt(pk, vb, 1, m2)[0]
        s = orc.apply(orc.OP_ADD, ca, cb, L)[0]
        return orc.decrypt(sk, s, L)[0]

    (la, lma), (lb, lmb) = shard(a, ma, rank, world, L, 4), shard(b, mb, rank, world, L, 4)
    local = run(la, lma, lb, lmb)
    parts = gather_plaintexts(local)
    t = max_over_ranks(1.0 + rank)
    if rank == 0:
        full = run(a, ma, b, mb)
        ret["ok"] = bool(np.array_equal(np.concatenate(parts), full)) and t == float(world)
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_rank(oracle):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get("ok") is True
