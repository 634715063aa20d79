#!/usr/bin/env python3
"""BASELINE.json configs[4] on 1/2/4/8 GPUs: PTO-style stitch of 6 rectilinear 6000x4000 positions
x 3 exposure brackets -> spherical 16384x8192, as the reference can run it:

  stage A   per position, `--synopsis hdr_merge --single 0` of its three brackets  (6 jobs)
  stage B   voronoi panorama of the six merged facets

One process per GPU (torchrun). Two plans:

--plan stripes (default): NO exchange between the stages. A rank's row band of the panorama sees a
  horizontal stripe of every position, so the rank merges exactly that stripe of each position itself
  (stage A through eu_render_rows on the stripe's rows, written in place into a full-size merged
  raster) and stitches its band from the six rasters, of which only the stripes are ever sampled.
  (PTO 'W' windows would save the full-size rasters, but the reference derives a window's vertical
  extent from the image WIDTH - environment.h:617,627 - which this back-end reproduces, so a window
  with a vertical offset does not mask like the image it was cut from.) Every rank holds all bracket
  rasters (broadcast once, the "source broadcast" of SURVEY 8e); stage B's kernels store the band
  straight into rank 0's frame over NVLink. Work per rank is ~1/N of both stages for any N (the
  positions plan cannot use more than six ranks for stage A).

--plan positions: the path sharded by independent units with two real exchange steps:
  1. rank 0 holds the 18 bracket rasters; each position's brackets are broadcast over NCCL
  2. stage A: positions are dealt round-robin to the ranks, each rank merges its own
  3. the merged facets are exchanged (NCCL broadcast from their owner) so every rank has all six
  4. stage B: every rank renders its row band of the panorama (eu_render_rows)
  5. the bands land on rank 0: by default the stage-B kernels store them straight into rank 0's frame
     over NVLink (eu_frame_*, step 4 and 5 are one kernel); --gather nccl keeps band buffers + gather
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
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--plan", default="stripes", choices=["stripes", "positions"])
    ap.add_argument("--bands", default="cost", choices=["cost", "equal"],
                    help="stripes plan: cost = band heights chosen so that every rank gets the same estimated "
                         "work (rows near the poles see no facet and cost little; SURVEY 8e 'balance by cost'); "
                         "equal = H/N rows each")
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
    windows_plan = a.plan == "stripes"
    H_pano, W_pano = 8192 // a.scale, 16384 // a.scale
    row0, row1 = eu_bands.band(H_pano, world, rank)

    def to_row(v):
        import math
        vext = math.tan(math.radians(50.0)) * h / w  # half the vertical extent of a position's image plane
        return (v / (2.0 * vext) + 0.5) * h - 0.5

    def cost_bands():
        """Band boundaries with equal estimated work. Per panorama row: stage B costs ~26 ps per pixel where
        a facet is in sight and a fraction of that elsewhere; the row also owns the rows of the six
        positions between its image and the next row's (stage A, ~35 ps per pixel)."""
        import math
        lat = ((np.arange(H_pano + 1)) / H_pano - 0.5) * math.pi
        lim = math.radians(89.9)
        fr = np.clip(to_row(np.tan(np.clip(lat, -lim, lim))), 0.0, float(h))
        corner = math.atan(math.tan(math.radians(50.0)) * h / w / math.cos(math.radians(50.0)))  # highest latitude in sight
        mid = 0.5 * (lat[:-1] + lat[1:])
        cost = W_pano * np.where(np.abs(mid) <= corner, 26.0, 6.0) + P * w * np.diff(fr) * 35.0
        return eu_bands.weighted_bands(cost, world)
    all_bands = eu_bands.bands(H_pano, world)
    if a.plan == "stripes" and a.bands == "cost" and world > 1 and a.gather == "peer":  # gather_bands wants equal bands
        all_bands = cost_bands()
    row0, row1 = all_bands[rank]

    def stripe(lat_rows):
        """Rows [r0, r1) of a position's 100-degree rectilinear image (pitch 0, roll 0) that the panorama
        rows [lat_rows) can touch: v = tan(lat) / cos(dlon), |dlon| <= 50 degrees, plus a margin for
        the bilinear window and the texel-centre convention."""
        import math
        la = ((lat_rows[0] + 0.5) / H_pano - 0.5) * math.pi
        lb = ((lat_rows[1] - 1 + 0.5) / H_pano - 0.5) * math.pi
        lim = math.radians(89.9)
        ta, tb = math.tan(max(-lim, min(lim, la))), math.tan(max(-lim, min(lim, lb)))
        c50 = math.cos(math.radians(50.0))
        vs = [ta, tb, ta / c50, tb / c50]
        r0 = int(math.floor(to_row(min(vs)))) - 8
        r1 = int(math.ceil(to_row(max(vs)))) + 9
        r0, r1 = max(0, min(h - 16, r0)), max(16, min(h, r1))
        return r0, max(r1, r0 + 16)
    wr0, wr1 = stripe((row0, row1)) if windows_plan else (0, h)

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
            if windows_plan or owner[p] == rank:
                d_br[(p, b)] = t
            else:
                del t
    e1.record()
    torch.cuda.synchronize()
    bcast_ms = dev_max(e0.elapsed_time(e1))

    eng = Engine(local)
    stream = torch.cuda.current_stream().cuda_stream

    # stage B's job description is needed first: in the stripes plan stage A renders straight into the
    # containers of stage B's sources (eu_source_reserve / eu_render_rows_pitched / eu_source_commit)
    fsb = [FacetSpec(None, "rectilinear", 100.0, yaw=60.0 * p, width=w, height=h, nchannels=3) for p in range(P)]
    jobb, algb = workloads.c5_stage_b_geometry(fsb, a.scale)
    stb = jobb.structs(eng.lib)
    reserved = {}

    # ---- 2. stage A on the owners ------------------------------------------------------
    merged = {}
    a_jobs = []
    stage_a_ms = 0.0
    for p in range(P):
        if not windows_plan and owner[p] != rank:
            continue
        fs = [FacetSpec(None, "rectilinear", 100.0, yaw=60.0 * p, eev=ev, width=w, height=h, nchannels=3) for ev in evs]
        job, alg = workloads.c5_stage_a_geometry(fs, w, h)
        st = job.structs(eng.lib)
        hs = eng.stage_device(job, [d_br[(p, b)].data_ptr() for b in range(B)], st, stream=stream)
        if p < 2:  # the first stagings grow the device pool: time the steady state
            eng.release(hs)
            hs = eng.stage_device(job, [d_br[(p, b)].data_ptr() for b in range(B)], st, stream=stream)
        stage_a_ms += sum(tm.render_ms for tm in eng.last_stage_timing)
        if windows_plan:
            # the merged image IS stage B's source: rows [wr0, wr1) are rendered into its container (the
            # other rows are never sampled by this rank's band)
            reserved[p] = eng.reserve(stb[1][p], stb[2])
            out = None
        else:
            out = torch.empty((h, w, 3), dtype=torch.float32, device=dev)
            merged[p] = out
        a_jobs.append((p, job, st, hs, out))

    def run_a():
        for p, job, st, hs, out in a_jobs:
            if windows_plan:
                hnd, core, pitch = reserved[p]
                eng.render_rows_pitched(job, hs, st, wr0, wr1, core + wr0 * pitch * 4, pitch, stream, timed=False)
            else:
                eng.render_rows(job, hs, st, 0, h, out.data_ptr(), stream, timed=False)
    run_a()
    a_ms = timed(run_a, a.steps)
    for p, job, st, hs, out in a_jobs:
        eng.release(hs)
    d_br.clear()

    # ---- 3. merged facets to everyone ----------------------------------------------------
    barrier()
    e0.record()
    for p in range(P):
        if windows_plan:
            break  # every rank has merged its own stripes: nothing to exchange
        if p not in merged:
            merged[p] = torch.empty((h, w, 3), dtype=torch.float32, device=dev)
        if world > 1:
            dist.broadcast(merged[p], owner[p])
    e1.record()
    torch.cuda.synchronize()
    xchg_ms = dev_max(e0.elapsed_time(e1))

    # ---- 4. stage B: row bands ------------------------------------------------------------
    H, W = stb[0].height, stb[0].width
    if windows_plan:  # the sources are in place: brace them
        from envutil_b200 import capi
        hsb = (capi.SourceH * P)()
        stage_b_ms = 0.0
        for p in range(P):
            stage_b_ms += eng.commit(reserved[p][0], stb[1][p], stb[2], stream).render_ms
            hsb[p] = reserved[p][0]
    else:
        hsb = eng.stage_device(jobb, [merged[p].data_ptr() for p in range(P)], stb, stream=stream)
        eng.release(hsb)  # steady state, as above
        hsb = eng.stage_device(jobb, [merged[p].data_ptr() for p in range(P)], stb, stream=stream)
        stage_b_ms = sum(tm.render_ms for tm in eng.last_stage_timing)
    assert (H, W) == (H_pano, W_pano)
    peer = None
    if world > 1 and a.gather == "peer":
        peer = eu_bands.PeerFrame(eng.lib, dist, H, W, 3, rank, world)
        out_ptr = peer.band_ptr(row0)
    else:
        band = torch.empty((row1 - row0, W, 3), dtype=torch.float32, device=dev)
        out_ptr = band.data_ptr()

    def run_b():
        eng.render_rows(jobb, hsb, stb, row0, row1, out_ptr, stream, timed=False)
    run_b()
    b_ms = timed(run_b, a.steps)

    # ---- 5. gather ----------------------------------------------------------------------
    gather_ms = 0.0
    if peer is not None:
        barrier()
        full = peer.as_tensor() if rank == 0 else None
    else:
        full = band
    if world > 1 and peer is None:
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
            "n_gpus": world, "steps": a.steps, "plan": a.plan,
            "stripe_rows_rank0": [wr0, wr1], "bands": [list(b) for b in all_bands],
            "stage_a": {"ms": a_ms, "mpix": mpix_a, "mpix_s": mpix_a / (a_ms * 1e-3),
                        "algorithmic_gbs_per_gpu": (P * w * h * 12 * 4 / world) / (a_ms * 1e-3) / 1e9,
                        "positions_per_rank": [P if windows_plan else owner.count(r) for r in range(world)],
                        "rows_per_position_rank0": wr1 - wr0, "staging_ms": stage_a_ms},
            "stage_b": {"ms": b_ms, "mpix": mpix_b, "mpix_s": mpix_b / (b_ms * 1e-3),
                        "algorithmic_gbs_per_gpu": (algb / world) / (b_ms * 1e-3) / 1e9,
                        "frac_measured_peak": (algb / world) / (b_ms * 1e-3) / 1e9 / peak,
                        "partition": "row bands", "staging_ms": stage_b_ms},
            "exchange": {"brackets_broadcast_ms": bcast_ms, "brackets_bytes": P * B * w * h * 12,
                         "merged_broadcast_ms": xchg_ms, "merged_bytes": P * w * h * 12,
                         "band_gather_ms": gather_ms, "band_bytes": W * H * 12,
                         "band_gather": ("peer stores into rank 0's frame inside stage B's kernels" if peer is not None
                                         else "NCCL gather of band buffers"),
                         "note": "bracket broadcast includes rank 0's pageable H2D copy of every raster"},
            "pipeline_ms_excluding_synthesis": a_ms + xchg_ms + stage_b_ms + b_ms + gather_ms,
            "checksum": checksum, "covered_fraction": covered, "synth_s": synth_s,
        }
        print(json.dumps(line))
    if peer is not None:
        full = None
        barrier()
        if rank != 0:
            peer.close()
        barrier()
        if rank == 0:
            peer.close()
    eng.release(hsb)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
