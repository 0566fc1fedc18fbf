"""CPU, world_size 2, gloo: the N > 1 path of bench.py — rank 0 owns the input and broadcasts it once,
blocks are independent so each rank parses its own contiguous shard with no data-path collective, and the
per-rank results concatenate to exactly the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BLOCK = 1 << 17


def shard_range(n_blocks: int, world: int, rank: int):
    """Contiguous block ranges, sizes differing by at most one (SURVEY.md 8e partitioning)."""
    base, extra = divmod(n_blocks, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _worker(rank, world, port, n_bytes, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as oracle
    from tests import datagen
    buf = torch.zeros(n_bytes, dtype=torch.uint8)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(datagen.mixed_corpus(n_bytes, seed=77)), dtype=torch.uint8))
    dist.broadcast(buf, 0)                                   # the only exchange on the path
    data = buf.numpy().tobytes()
    n_blocks = (n_bytes + BLOCK - 1) // BLOCK
    lo, hi = shard_range(n_blocks, world, rank)
    counts = torch.zeros(n_blocks, dtype=torch.int64)
    digest = torch.zeros(n_blocks, dtype=torch.int64)
    for b in range(lo, hi):
        s = oracle.model_block(data[b * BLOCK:(b + 1) * BLOCK], 3)
        counts[b] = len(s)
        digest[b] = int(s.astype(np.uint64).sum() % (1 << 62))
    dist.all_reduce(counts)                                  # test-only: gather the shards for the check
    dist.all_reduce(digest)
    if rank == 0:
        q.put((counts.tolist(), digest.tolist()))
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 8, 1617, 103488):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_ranks_equal_one(oracle):
    from tests import datagen
    n_bytes = 7 * BLOCK + 12345
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_bytes, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, digest = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = datagen.mixed_corpus(n_bytes, seed=77)
    for b in range(8):
        s = oracle.model_block(data[b * BLOCK:(b + 1) * BLOCK], 3)
        assert counts[b] == len(s)
        assert digest[b] == int(s.astype(np.uint64).sum() % (1 << 62))
