"""N > 1 host logic on CPU: the row-band partition and the band gather, run with two gloo ranks.
The per-band renderer here is the oracle (this is a test of the partition/gather plumbing, not
of the kernels); the GPU test test_row_bands_tile_the_full_render covers eu_render_rows."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import harness
import jobs
from envutil_b200 import bands


def test_band_partition_properties():
    for h in (1, 7, 8, 1080, 4096, 24576):
        for w in (1, 2, 3, 4, 8):
            bs = bands.bands(h, w)
            assert bs[0][0] == 0 and bs[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(bs, bs[1:]))
            sizes = [b[1] - b[0] for b in bs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        bands.band(10, 2, 2)


def test_weighted_bands():
    import json
    import math
    for world in (1, 2, 3, 8):  # uniform cost = the plain partition (up to where the odd row goes)
        wb = bands.weighted_bands(np.ones(4096), world)
        assert wb[0][0] == 0 and wb[-1][1] == 4096 and all(a[1] == b[0] for a, b in zip(wb, wb[1:]))
        assert max(b - a for a, b in wb) - min(b - a for a, b in wb) <= 1
    wb = bands.weighted_bands([0, 0, 0, 10, 0, 0, 0, 0], 4)  # all the work in one row: still one row each
    assert all(b > a for a, b in wb) and wb[-1][1] == 8
    with pytest.raises(ValueError):
        bands.weighted_bands([1.0, 1.0], 3)
    # the cost model of tools/bench_c5_multi.py (stage B per covered row + stage A rows it owns)
    # reproduces the bands recorded with the measured runs
    H, W, h, w, P = 8192, 16384, 4000, 6000, 6
    vext = math.tan(math.radians(50.0)) * h / w
    to_row = lambda v: (v / (2.0 * vext) + 0.5) * h - 0.5
    lat = (np.arange(H + 1) / H - 0.5) * math.pi
    lim = math.radians(89.9)
    fr = np.clip(to_row(np.tan(np.clip(lat, -lim, lim))), 0.0, float(h))
    corner = math.atan(math.tan(math.radians(50.0)) * h / w / math.cos(math.radians(50.0)))
    mid = 0.5 * (lat[:-1] + lat[1:])
    cost = W * np.where(np.abs(mid) <= corner, 26.0, 6.0) + P * w * np.diff(fr) * 35.0
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for world in (2, 4, 8):
        rec = json.load(open(os.path.join(root, "profiles", "r01g_c5_pipeline_n%d.json" % world)))["bands"]
        assert [list(b) for b in bands.weighted_bands(cost, world)] == rec


def _worker(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        job = jobs.JOBS[name]
        h = job.structs()[0].height
        r0, r1 = bands.band(h, world, rank)
        local = torch.from_numpy(harness.oracle_render(job, rows=(r0, r1), threads=1))
        full = bands.gather_bands(local, h, world, rank, dist)
        if rank == 0:
            q.put(full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["voronoi4_sph_d1", "ll_cube_d1"])  # 128 rows; 288 rows (6 faces)
def test_two_rank_band_render_equals_single(name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(full, harness.oracle_render(jobs.JOBS[name]))


def test_three_way_ragged_gather_single_process():
    """height % world != 0 with a fake in-process 'dist' (gather semantics only)."""
    class FakeDist:
        def __init__(self):
            self.parts = {}

        def gather(self, t, out, dst=0):
            self.parts[len(self.parts)] = t.clone()
            if out is not None:
                for i in range(len(out)):
                    out[i].copy_(self.parts[i])

    h, world = 10, 3
    img = torch.arange(h * 4, dtype=torch.float32).reshape(h, 4)
    fd = FakeDist()
    res = None
    for rank in (1, 2, 0):  # rank 0 last so that all parts are present
        r0, r1 = bands.band(h, world, rank)
        fd.parts[rank] = torch.zeros(4, 4)
        fd.parts[rank][: r1 - r0] = img[r0:r1]
    r0, r1 = bands.band(h, world, 0)

    class D:
        @staticmethod
        def gather(t, out, dst=0):
            for i in range(world):
                out[i].copy_(fd.parts[i])
    res = bands.gather_bands(img[r0:r1], h, world, 0, D)
    assert torch.equal(res, img)
