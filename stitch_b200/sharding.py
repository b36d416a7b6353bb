"""Read sharding across GPUs (the reference treats reads as independent,
fg-stitch-cli/src/commands/align.rs:345-379): contiguous blocks of the batch per rank, the contig table
replicated, results gathered in input order.  There is no data-path collective: the only communication
is the final ordered gather of the per-read results (and the timing reduction in bench.py)."""


def block_range(n_items: int, rank: int, world: int):
    """Contiguous block [lo, hi) of `n_items` owned by `rank` (blocks differ by at most one item)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_in_order(local_items, group=None):
    """All ranks' per-read results concatenated in rank (= input) order, on every rank."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(local_items)
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, list(local_items), group=group)
    return [x for p in parts for x in p]
