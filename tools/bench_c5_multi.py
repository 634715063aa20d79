#!/usr/bin/env python3
"""BASELINE.json configs[4] on 1/2/4/8 GPUs: PTO-style stitch of 6 rectilinear 6000x4000 positions
x 3 exposure brackets -> spherical 16384x8192, as the reference can run it:

  stage A   per position, `--synopsis hdr_merge --single 0` of its three brackets  (6 jobs)
  stage B   voronoi panorama of the six merged facets

One process per GPU (torchrun). The path shards by independent units with two real exchange steps:
  1. rank 0 holds the 18 bracket rasters; each position's brackets are broadcast over NCCL
  2. stage A: positions are dealt round-robin to the ranks, each rank merges its own
  3. the merged facets are exchanged (NCCL broadcast from their owner) so every rank has all six
  4. stage B: every rank renders its row band of the panorama (eu_render_rows)
  5. the bands are gathered on rank 0
Every phase is timed on the device (CUDA events, max over ranks); one JSON line on rank 0.

  python tools/bench_c5_multi.py [--scale 1] [--steps 5]
  python -m torch.distributed.run --nproc-per-node N ... tools/bench_c5_multi.py
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from envutil_b200 import bands as eu_bands  # noqa: E402
from envutil_b200 import workloads  # noqa: E402
from envutil_b200.engine import Engine  # noqa: E402
from envutil_b200.job import FacetSpec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5, help="timed repetitions of each render phase")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def dev_max(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps=1):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return dev_max(e0.elapsed_time(e1) / reps)

    P, B = 6, 3
    w, h = 6000 // a.scale, 4000 // a.scale
    t0 = time.time()
    if rank == 0:
        specs = workloads.c5_facets(a.scale)
        host = [torch.from_numpy(f.image) for f in specs]
    synth_s = time.time() - t0
    evs = workloads.C5_BRACKETS
    owner = [p % world for p in range(P)]

    # ---- 1. brackets -> owners (NCCL broadcast per raster) --------------------------------
    d_br = {}
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for p in range(P):
        for b in range(B):
            if rank == 0:
                t = host[p * B + b].to(dev, non_blocking=False)
            else:
                t = torch.empty((h, w, 3), dtype=torch.float32, device=dev)
            if world > 1:
                dist.broadcast(t, 0)
            if owner[p] == rank:
                d_br[(p, b)] = t
            else:
                del t
    e1.record()
    torch.cuda.synchronize()
    bcast_ms = dev_max(e0.elapsed_time(e1))

    eng = Engine(local)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- 2. stage A on the owners ------------------------------------------------------
    merged = {}
    a_jobs = []
    for p in range(P):
        if owner[p] != rank:
            continue
        fs = [FacetSpec(None, "rectilinear", 100.0, yaw=60.0 * p, eev=ev, width=w, height=h, nchannels=3) for ev in evs]
        job, alg = workloads.c5_stage_a_geometry(fs, w, h)
        st = job.structs(eng.lib)
        hs = eng.stage_device(job, [d_br[(p, b)].data_ptr() for b in range(B)], st, stream=stream)
        out = torch.empty((h, w, 3), dtype=torch.float32, device=dev)
        a_jobs.append((job, st, hs, out))
        merged[p] = out

    def run_a():
        for job, st, hs, out in a_jobs:
            eng.render_rows(job, hs, st, 0, h, out.data_ptr(), stream, timed=False)
    run_a()
    a_ms = timed(run_a, a.steps)
    for job, st, hs, out in a_jobs:
        eng.release(hs)
    d_br.clear()

    # ---- 3. merged facets to everyone ----------------------------------------------------
    barrier()
    e0.record()
    for p in range(P):
        if p not in merged:
            merged[p] = torch.empty((h, w, 3), dtype=torch.float32, device=dev)
        if world > 1:
            dist.broadcast(merged[p], owner[p])
    e1.record()
    torch.cuda.synchronize()
    xchg_ms = dev_max(e0.elapsed_time(e1))

    # ---- 4. stage B: row bands ------------------------------------------------------------
    fsb = [FacetSpec(None, "rectilinear", 100.0, yaw=60.0 * p, width=w, height=h, nchannels=3) for p in range(P)]
    jobb, algb = workloads.c5_stage_b_geometry(fsb, a.scale)
    stb = jobb.structs(eng.lib)
    H, W = stb[0].height, stb[0].width
    t_stage = time.perf_counter()
    hsb = eng.stage_device(jobb, [merged[p].data_ptr() for p in range(P)], stb, stream=stream)
    stage_b_ms = sum(tm.render_ms for tm in eng.last_stage_timing)
    row0, row1 = eu_bands.band(H, world, rank)
    band = torch.empty((row1 - row0, W, 3), dtype=torch.float32, device=dev)

    def run_b():
        eng.render_rows(jobb, hsb, stb, row0, row1, band.data_ptr(), stream, timed=False)
    run_b()
    b_ms = timed(run_b, a.steps)

    # ---- 5. gather ----------------------------------------------------------------------
    gather_ms = 0.0
    full = band
    if world > 1:
        full = eu_bands.gather_bands(band, H, world, rank, dist)  # sets up the channels
        barrier()
        e0.record()
        full = eu_bands.gather_bands(band, H, world, rank, dist)
        e1.record()
        torch.cuda.synchronize()
        gather_ms = e0.elapsed_time(e1)
    if rank == 0:
        checksum = float(full[::61, ::67].double().sum().item())
        covered = float((full[::16, ::16].sum(dim=2) > 0).double().mean().item())
        mpix_a, mpix_b = P * w * h / 1e6, W * H / 1e6
        peak = 6537.0
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peak = float(json.load(open(pk))["hbm_gbs"])
        line = {
            "workload": "C5: 6 rectilinear %dx%d positions x 3 brackets -> hdr_merge per position (--single 0) -> "
                        "voronoi panorama spherical %dx%d" % (w, h, W, H),
            "n_gpus": world, "steps": a.steps,
            "stage_a": {"ms": a_ms, "mpix": mpix_a, "mpix_s": mpix_a / (a_ms * 1e-3),
                        "algorithmic_gbs_per_gpu": (P * w * h * 12 * 4 / world) / (a_ms * 1e-3) / 1e9,
                        "positions_per_rank": [owner.count(r) for r in range(world)]},
            "stage_b": {"ms": b_ms, "mpix": mpix_b, "mpix_s": mpix_b / (b_ms * 1e-3),
                        "algorithmic_gbs_per_gpu": (algb / world) / (b_ms * 1e-3) / 1e9,
                        "frac_measured_peak": (algb / world) / (b_ms * 1e-3) / 1e9 / peak,
                        "partition": "row bands", "staging_ms": stage_b_ms},
            "exchange": {"brackets_broadcast_ms": bcast_ms, "brackets_bytes": P * B * w * h * 12,
                         "merged_broadcast_ms": xchg_ms, "merged_bytes": P * w * h * 12,
                         "band_gather_ms": gather_ms, "band_bytes": W * H * 12,
                         "note": "bracket broadcast includes rank 0's pageable H2D copy of every raster"},
            "pipeline_ms_excluding_synthesis": a_ms + xchg_ms + stage_b_ms + b_ms + gather_ms,
            "checksum": checksum, "covered_fraction": covered, "synth_s": synth_s,
        }
        print(json.dumps(line))
    eng.release(hsb)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
