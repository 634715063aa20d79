"""Row-band partition of a target over the ranks of one box: contiguous bands,
sizes differing by at most one row, gathered in rank order. Pure host logic (no GPU)."""


def band(height, world, rank):
    """Rows [row0, row1) rendered by `rank` of `world`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(height, world)
    row0 = rank * base + min(rank, extra)
    return row0, row0 + base + (1 if rank < extra else 0)


def bands(height, world):
    return [band(height, world, r) for r in range(world)]


def gather_bands(local, height, world, rank, dist, dst=0):
    """Assemble the full frame on `dst` from per-rank band tensors (torch.distributed; works with
    gloo on CPU tensors and nccl on CUDA tensors). Bands may be ragged (height % world != 0):
    they are padded to the tallest band for the collective and trimmed afterwards."""
    import torch
    tallest = max(r1 - r0 for r0, r1 in bands(height, world))
    pad = torch.zeros((tallest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, out, dst=dst)
    if rank != dst:
        return None
    parts = [out[r][: b[1] - b[0]] for r, b in enumerate(bands(height, world))]
    return torch.cat(parts, dim=0)
