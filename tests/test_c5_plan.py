"""Host logic of the configs[4] pipeline (envutil_b200/c5.py): the row-band partition and the rectangles of the
position rasters a band can sample. CPU only (numpy; one gloo test with two ranks)."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

from envutil_b200 import c5

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sampled_texels(row0, row1, w, h, W, H):
    """Texel windows (position, column, row of the window's top-left texel) that stage B's bilinear evaluator
    reads for the panorama rows [row0, row1), from the projection formulas in float64."""
    ex, ey = c5._extent(w, h)
    lat = ((np.arange(row0, row1) + 0.5) / H - 0.5) * math.pi
    lon = ((np.arange(W) + 0.5) / W - 0.5) * 2.0 * math.pi
    LON, LAT = np.meshgrid(lon, lat)
    out = []
    for p in range(c5.POSITIONS):
        yaw = math.radians(c5.YAW_STEP_DEG * p)
        d = (LON - yaw + math.pi) % (2.0 * math.pi) - math.pi
        near = np.abs(d) <= math.radians(c5.YAW_STEP_DEG / 2.0) + 1e-9  # the voronoi winner is the nearest in longitude
        u = np.tan(d)
        v = np.tan(LAT) / np.cos(d)
        hit = near & (np.abs(v) <= ey) & (np.abs(u) <= ex)
        cx = c5.col_of(u[hit], w, h)
        cy = c5.row_of(v[hit], w, h)
        out.append((np.floor(cx).astype(int), np.floor(cy).astype(int)))
    return out


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_rectangles_cover_what_the_band_samples(world):
    """Every 2 x 2 window stage B can read from a position lies inside one of the band's rectangles (clipped to the
    raster: beyond it the brace, a copy of the edge texels, is read), with a texel to spare."""
    scale = 10
    (w, h), (W, H) = c5.sizes(scale)
    bands = c5.plan_bands(world, scale)
    assert bands[0][0] == 0 and bands[-1][1] == H and all(a[1] == b[0] for a, b in zip(bands[:-1], bands[1:]))
    for row0, row1 in bands:
        rects = c5.rects_for_band(row0, row1, w, h, H)
        for r0, r1, c0, c1 in rects:
            assert c0 % 32 == 0 and 0 <= r0 < r1 <= h and 0 <= c0 < c1 <= w
        for ix, iy in _sampled_texels(row0, row1, w, h, W, H):
            if ix.size == 0:
                continue
            covered = np.zeros(ix.shape, dtype=bool)
            x0, x1 = np.clip(ix, 0, w - 1), np.clip(ix + 1, 0, w - 1)
            y0, y1 = np.clip(iy, 0, h - 1), np.clip(iy + 1, 0, h - 1)
            for r0, r1, c0, c1 in rects:
                # one texel to spare on every side that is not the raster's edge
                lo_r, hi_r = (r0 if r0 == 0 else r0 + 1), (r1 if r1 == h else r1 - 1)
                lo_c, hi_c = (c0 if c0 == 0 else c0 + 1), (c1 if c1 == w else c1 - 1)
                covered |= (x0 >= lo_c) & (x1 < hi_c) & (y0 >= lo_r) & (y1 < hi_r)
            # windows that straddle two strips are covered by the union: check texel by texel for the rest
            rest = ~covered
            if rest.any():
                ok = np.ones(int(rest.sum()), dtype=bool)
                for tx, ty in ((x0, y0), (x1, y0), (x0, y1), (x1, y1)):
                    inside = np.zeros(ok.shape, dtype=bool)
                    for r0, r1, c0, c1 in rects:
                        inside |= (tx[rest] >= c0) & (tx[rest] < c1) & (ty[rest] >= r0) & (ty[rest] < r1)
                    ok &= inside
                assert ok.all(), (world, row0, row1)


def test_needed_region_is_half_the_raster_at_full_size():
    (w, h), (W, H) = c5.sizes(1)
    c0, c1 = c5.needed_columns(w, h)
    assert c0 % 32 == 0 and 0.47 < (c1 - c0) / w < 0.51
    # one rank: the rectangles are the needed columns over all rows; eight ranks overlap by less than a fifth
    one = c5.stage_a_pixels(c5.rects_for_band(0, H, w, h, H))
    assert one == h * (c1 - c0)
    eight = sum(c5.stage_a_pixels(c5.rects_for_band(a, b, w, h, H)) for a, b in c5.plan_bands(8, 1))
    assert one <= eight < 1.2 * one


def test_cost_bands_are_balanced():
    (w, h), (W, H) = c5.sizes(1)
    cost = c5.row_costs(w, h, W, H)
    for world in (2, 4, 8):
        per = [cost[a:b].sum() for a, b in c5.plan_bands(world, 1)]
        assert max(per) < 1.1 * (sum(per) / world)


def test_transfer_bytes_per_rank_are_even_at_eight_ranks():
    """bands sized by device cost also spread the PCIe traffic - the polar ranks download most and
    upload next to nothing - so a rank's H2D + D2H bytes stay within 469 ... 594 MB of the 544 MB mean at N = 8, and the
    rectangles of all ranks together are 0.529 of the rasters."""
    (w, h), (W, H) = c5.sizes(1)
    bands = c5.plan_bands(8, 1)
    sums, merged = [], 0
    for r0, r1 in bands:
        rects = c5.rects_for_band(r0, r1, w, h, H)
        h2d = c5.stage_a_pixels(rects) * c5.POSITIONS * c5.BRACKETS * 12
        merged += c5.stage_a_pixels(rects)
        sums.append(h2d + (r1 - r0) * W * 12)
    assert 460e6 < min(sums) and max(sums) < 600e6
    assert abs(merged / (w * h) - 0.529) < 0.01


def test_windowed_synthesis_equals_the_whole_raster():
    from envutil_b200 import synth, workloads
    fs = workloads.c5_facets(scale=20, positions=2)
    (w, h) = (6000 // 20, 4000 // 20)
    rect = (13, 150, 32, 201)
    for p in range(2):
        for b, img in enumerate(c5.synth_rect(p, rect, w, h)):
            assert np.array_equal(img, fs[p * 3 + b].image[rect[0]:rect[1], rect[2]:rect[3]])


def _gloo_worker(rank, world, port, path, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        (w, h), (W, H) = c5.sizes(40)
        bands = c5.plan_bands(world, 40)
        got = [None] * world
        dist.all_gather_object(got, bands)
        assert all(g == bands for g in got)
        if rank == 0:
            np.lib.format.open_memmap(path, mode="w+", dtype=np.float32, shape=(H, W, 3)).flush()
        dist.barrier()
        frame = np.load(path, mmap_mode="r+")  # ONE frame in shared host memory, every rank writes its own band
        r0, r1 = bands[rank]
        frame[r0:r1] = np.arange(r0, r1, dtype=np.float32)[:, None, None] + 1.0
        frame.flush()
        dist.barrier()
        if rank == 0:
            whole = np.load(path, mmap_mode="r")
            q.put((bands, bool(np.array_equal(whole[:, 0, 0], np.arange(H, dtype=np.float32) + 1.0))))
    finally:
        dist.destroy_process_group()


def test_ranks_agree_on_the_partition_gloo():
    """Two processes (gloo): every rank derives the same bands from the same arguments, the bands tile the
    panorama, and the frame assembled from per-rank bands in shared host memory is complete."""
    import tempfile
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + ((os.getpid() + 977) % 2000)
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, os.path.join(d, "frame.npy"), q)) for r in range(2)]
        for p in procs:
            p.start()
        bands, complete = q.get(timeout=120)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    assert complete and bands[0][0] == 0 and bands[-1][1] == c5.sizes(40)[1][1]
