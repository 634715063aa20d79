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


def weighted_bands(row_cost, world):
    """Contiguous bands with (nearly) equal summed cost - SURVEY 8e: 'optionally balance by cost' where
    the per-row work varies (rows of a panorama that no facet covers, fisheye corners ...). row_cost:
    one non-negative number per row. Every band has at least one row; returns [(row0, row1)] * world."""
    import numpy as np
    cost = np.asarray(row_cost, dtype=np.float64)
    height = int(cost.size)
    if world < 1 or height < world:
        raise ValueError("need at least one row per rank")
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    edges = [0] + [int(np.searchsorted(cum, cum[-1] * k / world)) for k in range(1, world)] + [height]
    for k in range(1, world):            # strictly increasing, leaving a row for every later band
        edges[k] = min(max(edges[k], edges[k - 1] + 1), height - (world - k))
    return [(edges[k], edges[k + 1]) for k in range(world)]


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


def gather_ragged(local, all_bands, rank, dist, dst=0):
    """gather_bands for bands of ANY heights (cost-balanced partitions): all_bands = [(row0, row1)] per rank."""
    import torch
    tallest = max(r1 - r0 for r0, r1 in all_bands)
    pad = torch.zeros((tallest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in all_bands] if rank == dst else None
    dist.gather(pad, out, dst=dst)
    if rank != dst:
        return None
    return torch.cat([out[r][: b[1] - b[0]] for r, b in enumerate(all_bands)], dim=0)


class PeerFrame:
    """One frame in the HBM of rank `dst` that every rank renders its row band into (fused render +
    gather: the kernel's stores travel over NVLink, include/envutil_b200.h eu_frame_*). The owner
    allocates and exports, the handle travels through torch.distributed's object broadcast (any
    backend), the other ranks open it. band_ptr(row0) is the d_out to hand to eu_render_rows."""

    def __init__(self, lib, dist, height, width, nch, rank, world, dst=0):
        import ctypes as C
        from . import capi
        self.lib, self.rank, self.dst, self.row_bytes = lib, rank, dst, width * nch * 4
        self.height, self.width, self.nch = height, width, nch
        self.ptr = C.c_void_p()
        box = [None]
        if rank == dst:
            capi.check(lib.eu_frame_alloc(height * width * nch, C.byref(self.ptr)), lib)
            buf = C.create_string_buffer(64)
            capi.check(lib.eu_frame_export(self.ptr, buf), lib)
            box[0] = buf.raw
        if world > 1:
            dist.broadcast_object_list(box, src=dst)
        if rank != dst:
            capi.check(lib.eu_frame_open(box[0], C.byref(self.ptr)), lib)

    def band_ptr(self, row0):
        return self.ptr.value + row0 * self.row_bytes

    def as_tensor(self):
        """The whole frame as a CUDA tensor (owner only; valid until close())."""
        import torch
        assert self.rank == self.dst
        class _Mem:  # __cuda_array_interface__ carrier
            pass
        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (self.height, self.width, self.nch), "typestr": "<f4",
                                      "data": (self.ptr.value, False), "version": 3, "strides": None}
        return torch.as_tensor(m, device="cuda")

    def close(self):
        from . import capi
        if self.ptr.value:
            fn = self.lib.eu_frame_free if self.rank == self.dst else self.lib.eu_frame_close
            capi.check(fn(self.ptr), self.lib)
            self.ptr.value = None
