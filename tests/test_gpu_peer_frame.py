"""Fused render + gather: two ranks (one process per GPU) render their row bands straight into
rank 0's frame through eu_frame_* (peer stores over NVLink, no band buffers, no collective in the
data path); the assembled frame equals the oracle's bit for bit. Needs two GPUs - skipped on a
one-GPU box (the single-GPU part, a frame handle opened inside the exporting box, cannot be
tested in one process: CUDA IPC handles open in OTHER processes only)."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import harness
import jobs

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from envutil_b200 import bands
    from envutil_b200.engine import Engine
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    eng = Engine(rank)
    try:
        job = jobs.JOBS[name]
        st = job.structs(eng.lib)
        t = st[0]
        hs = eng.stage(job, st)
        frame = bands.PeerFrame(eng.lib, dist, t.height, t.width, t.nchannels, rank, world)
        r0, r1 = bands.band(t.height, world, rank)
        eng.render_rows(job, hs, st, r0, r1, frame.band_ptr(r0), 0, timed=True)  # timed: returns after the kernel
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            q.put(frame.as_tensor().cpu().numpy())
        dist.barrier()
        if rank != 0:
            frame.close()
        dist.barrier()
        if rank == 0:
            frame.close()
        eng.release(hs)
    finally:
        eng.close()
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["cm_sph_d3", "voronoi4_sph_d1", "ll_cube_d1", "ll_fish_d1_tw4"])
def test_two_gpus_render_bands_into_one_frame(name):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(full, harness.oracle_render(jobs.JOBS[name]))
