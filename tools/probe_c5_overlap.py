#!/usr/bin/env python3
"""Do the upload of step n+1 and the download of step n of the configs[4] pipeline overlap on this box?

    python tools/probe_c5_overlap.py                                   (one GPU)
    python -m torch.distributed.run --nproc-per-node N ... tools/probe_c5_overlap.py

Every rank builds its share of the pipeline (envutil_b200/c5.py) and times, all ranks at the same moment, wall clock:
  U        its H2D (eu_source_write_rect of its rectangles + brace), alone
  D        its band's D2H, alone - into (a) the shared page-locked frame (/dev/shm + cudaHostRegister, what bench.py
           uses) and (b) a cudaHostAlloc buffer of the same size
  U || D   both at once on two streams (no kernels), for (a) and (b)
  flat U   the same bytes as ONE contiguous copy from a page-locked buffer
Diagnostic tool, not a bench: prints one JSON line per rank."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from envutil_b200 import c5
from envutil_b200.engine import Engine

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
balance = sys.argv[1] if len(sys.argv) > 1 else "cost"
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


(w, h), (W, H) = c5.sizes(1)
eng = Engine(local)
tag = os.environ.get("MASTER_PORT", "0") + "_" + str(os.getppid() if world > 1 else os.getpid())
shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
path = os.path.join(shm, "eu_probe_frame_%s.f32" % tag)
if rank == 0:
    np.memmap(path, dtype=np.float32, mode="w+", shape=(H, W, 3)).flush()
barrier()
frame_np = np.memmap(path, dtype=np.float32, mode="r+", shape=(H, W, 3))
frame_t = torch.from_numpy(frame_np)
pl = c5.Pipeline(eng, torch, rank, world, 1, balance=balance, host_frame=frame_t)
band_shm = frame_t[pl.row0:pl.row1]
band_shm.zero_()
rt = torch.cuda.cudart()
rt.cudaHostRegister(band_shm.data_ptr(), band_shm.numel() * 4, 0)
band_pin = torch.empty(band_shm.shape, dtype=torch.float32).pin_memory()
flat_host = torch.empty(pl.h2d_bytes // 4, dtype=torch.float32).pin_memory()
flat_dev = torch.empty(pl.h2d_bytes // 4, dtype=torch.float32, device="cuda")
copy_stream = torch.cuda.Stream()
main = torch.cuda.current_stream()


def timed(fn, reps=4):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def up():
    pl.upload()


def down(dst):
    def f():
        with torch.cuda.stream(copy_stream):
            dst.copy_(pl.d_band[0], non_blocking=True)
    return f


def both(dst):
    def f():
        with torch.cuda.stream(copy_stream):
            dst.copy_(pl.d_band[0], non_blocking=True)
        pl.upload()
    return f


def flat_up():
    flat_dev.copy_(flat_host, non_blocking=True)


def flat_both(dst):
    def f():
        with torch.cuda.stream(copy_stream):
            dst.copy_(pl.d_band[0], non_blocking=True)
        flat_dev.copy_(flat_host, non_blocking=True)
    return f


res = {"rank": rank, "world": world, "balance": balance, "rows": [pl.row0, pl.row1], "h2d_mb": pl.h2d_bytes / 1e6,
       "d2h_mb": pl.d2h_bytes / 1e6}
res["U_ms"] = timed(up)
res["D_shm_ms"] = timed(down(band_shm))
res["D_pin_ms"] = timed(down(band_pin))
res["UD_shm_ms"] = timed(both(band_shm))
res["UD_pin_ms"] = timed(both(band_pin))
res["flatU_ms"] = timed(flat_up)
res["flatUD_shm_ms"] = timed(flat_both(band_shm))
res["flatUD_pin_ms"] = timed(flat_both(band_pin))
# the real e2e loop (bench.py's), 6 steps
for _ in range(2):
    pl.step_e2e()
pl.finish()
barrier()
t0 = time.perf_counter()
for _ in range(6):
    pl.step_e2e()
pl.finish()
barrier()
res["e2e_ms_per_step"] = (time.perf_counter() - t0) / 6 * 1e3
got = [None] * world
if world > 1:
    dist.all_gather_object(got, res)
else:
    got = [res]
if rank == 0:
    for r in got:
        print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)
barrier()
rt.cudaHostUnregister(band_shm.data_ptr())
pl.close()
eng.close()
del frame_t, frame_np, band_shm
barrier()
if rank == 0:
    try:
        os.unlink(path)
    except OSError:
        pass
if world > 1:
    dist.destroy_process_group()
