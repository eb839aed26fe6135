"""Multi-rank host logic on CPU: gloo, world_size 2 (the GPU path uses NCCL)."""
import os
import sys

import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from localuf_b200 import dist as ldist
    from localuf_b200.decoders._base import Decoder
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        assert ldist.world() == (rank, world)
        assert ldist.allreduce_counts([rank + 1, 10]) == [3, 20]

        class Fake(Decoder):
            class _E:
                device = 0
            engine = _E()

        seen = {}

        def run_local(seed, lo, cnt):
            seen.update(seed=seed, lo=lo, cnt=cnt)
            return cnt // 10, cnt           # "failures", shots

        dec = Fake.__new__(Fake)
        total = dec._sharded_counts(run_local, 1001, seed=None if rank else 777)
        lo, cnt = ldist.shard(1001, rank, world)
        assert (seen['lo'], seen['cnt']) == (lo, cnt)
        assert seen['seed'] == 777              # rank 0's seed reaches every rank
        out.put((rank, total, lo, cnt))
    finally:
        dist.destroy_process_group()


def test_shard_covers_range():
    from localuf_b200.dist import shard
    for n in (0, 1, 7, 1000, 10**9 + 7):
        for w in (1, 2, 3, 8):
            parts = [shard(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(w - 1))


def test_two_rank_gloo_allreduce():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(out.get(timeout=5) for _ in range(2))
    assert res[0][1] == res[1][1] == 500 // 10 + 501 // 10
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 500, 500, 501)
