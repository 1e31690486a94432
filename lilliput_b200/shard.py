"""Host-side multi-GPU plumbing for the batch path: images are independent, so GPUs never exchange
pixel data.  A job over `n_total` images is sharded by image index (`i mod world`);
torch.distributed is used only for the barrier around timed regions and the max-over-ranks reduction
of the timings.  Works with the `nccl` backend on GPUs and `gloo` on CPU (tests)."""
from __future__ import annotations


def shard_indices(n_total: int, rank: int, world: int) -> list[int]:
    """Image indices rank `rank` of `world` processes: i mod world == rank (round robin keeps the
    per-rank byte volume balanced when image sizes vary with index, as in BASELINE config 5)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_total, world))


def corpus_seed(base: int, rank: int, per_rank: int) -> int:
    """Seed of the first image of rank's private corpus in the weak-scaling bench."""
    return base + rank * per_rank


def merge_sharded(results_per_rank: list[list], n_total: int) -> list:
    """Inverse of shard_indices: interleave per-rank result lists back into index order."""
    world = len(results_per_rank)
    out = [None] * n_total
    for r, res in enumerate(results_per_rank):
        idx = shard_indices(n_total, r, world)
        if len(res) != len(idx):
            raise ValueError("rank %d returned %d results for %d images" % (r, len(res), len(idx)))
        for i, v in zip(idx, res):
            out[i] = v
    return out


def max_over_ranks(values: list[float], dist=None, device="cpu") -> list[float]:
    """Element-wise MAX of `values` over all ranks (identity without a process group)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(values)
    import torch
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]
