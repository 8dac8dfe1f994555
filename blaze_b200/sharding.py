"""Host-side sharding helpers for the multi-GPU paths (one process per GPU, torch.distributed).

MSM shards by points with no data-path collective: rank g keeps points / scalars
[g N/G, (g+1) N/G) resident and the only exchange is an all-gather of the G result records
(144 or 96 bytes each), which any rank then sums (`MSMClient.combine_results` on the GPU).
"""


def shard_range(n, rank, world):
    """[first, first + count) of rank's contiguous shard; the remainder goes to the low ranks."""
    base, rem = divmod(n, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def gather_records(dist, record: bytes, device=None):
    """all-gather one small byte record per rank; returns the list in rank order.
    `dist` is torch.distributed (initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [bytes(record)]
    import torch
    t = torch.frombuffer(bytearray(record), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [bytes(o.cpu().numpy()) for o in out]
