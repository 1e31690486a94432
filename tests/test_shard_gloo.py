"""CPU: the N>1 plumbing of the batch path under world_size 2 with the gloo backend (no GPU):
index sharding covers every image exactly once, results merge back in order, and the timing
reduction takes the max over ranks."""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from lilliput_b200.shard import max_over_ranks, shard_indices
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = shard_indices(11, rank, world)
    # each rank "processes" its shard: result = index squared; elapsed differs per rank
    res = [i * i for i in idx]
    elapsed = [0.5 + rank, 2.0 - rank]
    dist.barrier()
    mx = max_over_ranks(elapsed, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    q.put((rank, idx, mx, gathered))
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_reduce():
    from lilliput_b200.shard import merge_sharded, shard_indices
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort()
    all_idx = sorted(out[0][1] + out[1][1])
    assert all_idx == list(range(11))                      # every image exactly once
    assert out[0][2] == out[1][2] == [1.5, 2.0]            # max over ranks, same on both
    merged = merge_sharded(out[0][3], 11)
    assert merged == [i * i for i in range(11)]            # back in index order
    assert shard_indices(11, 1, 2) == [1, 3, 5, 7, 9]


def test_shard_helpers_single_process():
    from lilliput_b200.shard import corpus_seed, max_over_ranks, shard_indices
    assert shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    assert shard_indices(0, 0, 4) == []
    assert sum(len(shard_indices(65536, r, 8)) for r in range(8)) == 65536
    assert corpus_seed(1000, 3, 4096) == 1000 + 3 * 4096
    assert max_over_ranks([1.0, 2.0]) == [1.0, 2.0]
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def test_library_block_sharding_is_contiguous_balanced_and_complete():
    """lp_shard_blocks (xbatch.cu, host only): the cut lp_multi_transform makes -- contiguous blocks, every item in
    exactly one, byte-balanced, degenerate inputs handled.  No device call."""
    import ctypes as C
    import os
    import numpy as np
    lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lilliput_b200", "liblilliput_b200.so"))
    lib.lp_shard_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(3)
    for n, parts in [(0, 4), (1, 8), (7, 8), (8, 8), (1000, 8), (4096, 3), (65536, 8)]:
        lens = rng.integers(1000, 30_000_000 if n < 5000 else 2_000_000, max(n, 1), dtype=np.uint64)[:n]
        a = np.ascontiguousarray(lens, dtype=np.uint64)
        first = np.full(parts + 1, -1, np.int32)
        lib.lp_shard_blocks(a.ctypes.data if n else None, n, parts, first.ctypes.data)
        assert first[0] == 0 and first[parts] == n and all(first[p] <= first[p + 1] for p in range(parts))
        if n >= 100 * parts:
            sums = [int(a[first[p]:first[p + 1]].sum()) for p in range(parts)]
            assert max(sums) <= 1.2 * (sum(sums) / parts) + int(a.max())
