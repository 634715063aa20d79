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
