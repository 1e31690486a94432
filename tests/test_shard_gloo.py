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
